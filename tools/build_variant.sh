#!/bin/bash
# Builds a variant of the library + oc2cns/oc2pmov with extra nvcc defines into build/var_<name>/{lib,bin} (for A/B runs on the GPU box).
# usage: tools/build_variant.sh <name> "<defines>"
set -e
cd "$(dirname "$0")/.."
N=$1; D=$2; V=build/var_$N
mkdir -p $V/obj $V/lib $V/bin
for f in necat_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 $D -c $f -o $V/obj/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/lib/libnecat_b200.so $V/obj/*.o -cudart static
for e in oc2cns oc2pmov; do
  gcc -O2 -g -std=gnu99 -o $V/bin/$e necat_b200/host/$e.c necat_b200/host/nb_host.c necat_b200/host/nb_helper.c -L$V/lib -lnecat_b200 -Wl,-rpath,'$ORIGIN/../lib' -lm
done
echo built $V
