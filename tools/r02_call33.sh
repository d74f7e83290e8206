#!/bin/bash
# Round 2, GPU call 33: where oc2cns / oc2pmov spend their time on the config-4-size set (5.6 Gbp, three volumes, eight
# partitions): NB_TRACE phases of one partition and of the first index volume.
mkdir -p gpurun_out
O=gpurun_out
W=/tmp/nbs; mkdir -p $W
timeout 300 python - <<PY
import sys, os, multiprocessing as mp
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import scale_volumes as S
G, D, P = 140_000_000, 40, 16
fa = ["$W/reads_%02d.fa" % w for w in range(P)]
with mp.Pool(P) as pool:
    pool.map(S._worker, [(G, 20261017, w, int(G * D) // P, 0, fa[w]) for w in range(P)])
open("$W/list", "w").write("".join(p + "\n" for p in fa))
PY
B=necat_b200/bin
PM="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CN="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
$B/oc2mkdb $W/wrk $W/list > /dev/null 2>&1
NB_TRACE=1 timeout 120 $B/oc2pmov $PM -t 16 $W/wrk 0 $W/c0.bin 2>&1 > /dev/null | grep -E "phase|nb_map_range" | cut -c1-300 | tail -12 | tee $O/c33_pmov_v0_trace.txt
timeout 120 $B/oc2pmov $PM -t 16 $W/wrk 1 $W/c1.bin > /dev/null 2>&1
timeout 120 $B/oc2pmov $PM -t 16 $W/wrk 2 $W/c2.bin > /dev/null 2>&1
cat $W/c0.bin $W/c1.bin $W/c2.bin > $W/cand.bin
timeout 60 $B/oc2pcan -t 16 $W/wrk $W/cand.bin > /dev/null 2>&1
( time NB_TRACE=1 timeout 120 $B/oc2cns $CN -t 16 $W/wrk $W/cand.bin $W/c.fa $W/r.fa -mn 6 8 ) 2>&1 > /dev/null | grep -E "phase|nb_consensus_job|real" | cut -c1-600 | tee $O/c33_cns_p6_trace.txt
( time NECAT_B200_HELPER=1 timeout 120 $B/oc2cns $CN -t 16 $W/wrk $W/cand.bin $W/c.fa $W/r.fa -mn 5 8 ) 2>&1 > /dev/null | grep real
( time NECAT_B200_HELPER=1 NB_TRACE=1 timeout 120 $B/oc2cns $CN -t 16 $W/wrk $W/cand.bin $W/c.fa $W/r.fa -mn 6 8 ) 2>&1 > /dev/null | grep -E "phase|real" | tee $O/c33_cns_p6_helper_trace.txt
rm -rf $W
