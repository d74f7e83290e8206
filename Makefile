# Builds libnecat_b200.so (sm_100a only) and the drop-in executables, in-tree.
NVCC  ?= /usr/local/cuda/bin/nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 -Xptxas -v
CSRC  := necat_b200/csrc
LIBD  := necat_b200/lib
BIND  := necat_b200/bin
OBJD  := build/obj
CU    := $(wildcard $(CSRC)/*.cu)
OBJ   := $(patsubst $(CSRC)/%.cu,$(OBJD)/%.o,$(CU))
HOSTC := necat_b200/host/nb_host.c
CC    := gcc
CFLAGS := -O2 -g -std=gnu99 -Wall

all: $(LIBD)/libnecat_b200.so $(BIND)/oc2pmov $(BIND)/oc2cns $(BIND)/oc2pcan $(BIND)/oc2mkdb $(BIND)/oc2ReorderCnsReads $(BIND)/nb_helperd \
     $(BIND)/fsa_rd_tools $(BIND)/pigz

$(OBJD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/necat_b200.h
	@mkdir -p $(OBJD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJD)/$*.ptxas.log || (cat $(OBJD)/$*.ptxas.log; false)

$(LIBD)/libnecat_b200.so: $(OBJ)
	@mkdir -p $(LIBD)
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

# ingest / reorder hosts: plain C + zlib, no CUDA library (they run where there is no GPU)
FASTX := necat_b200/host/nb_fastx.c
$(BIND)/oc2mkdb $(BIND)/oc2ReorderCnsReads: $(BIND)/%: necat_b200/host/%.c $(HOSTC) $(FASTX) necat_b200/host/nb_host.h necat_b200/host/nb_fastx.h
	@mkdir -p $(BIND)
	$(CC) $(CFLAGS) -O3 -o $@ $< $(HOSTC) $(FASTX) -lz -lpthread

# the two helpers necat.pl runs around a correction iteration: self-contained C + zlib
$(BIND)/fsa_rd_tools $(BIND)/pigz: $(BIND)/%: necat_b200/host/%.c
	@mkdir -p $(BIND)
	$(CC) $(CFLAGS) -O3 -o $@ $< -lz -lpthread

HELPER := necat_b200/host/nb_helper.c
$(BIND)/%: necat_b200/host/%.c $(HOSTC) $(HELPER) necat_b200/host/nb_host.h necat_b200/host/nb_helper.h $(LIBD)/libnecat_b200.so
	@mkdir -p $(BIND)
	$(CC) $(CFLAGS) -o $@ $< $(HOSTC) $(HELPER) -L$(LIBD) -lnecat_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lm -lpthread

# the resident helper (NECAT_B200_HELPER=1): the two tools as functions of one process that keeps the CUDA context
$(BIND)/nb_helperd: necat_b200/host/nb_helperd.c necat_b200/host/oc2pmov.c necat_b200/host/oc2cns.c $(HOSTC) $(HELPER) necat_b200/host/nb_host.h necat_b200/host/nb_helper.h $(LIBD)/libnecat_b200.so
	@mkdir -p $(BIND)
	$(CC) $(CFLAGS) -DNBH_AS_LIBRARY -o $@ necat_b200/host/nb_helperd.c necat_b200/host/oc2pmov.c necat_b200/host/oc2cns.c $(HOSTC) $(HELPER) -L$(LIBD) -lnecat_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lm -lpthread

clean:
	rm -rf build $(LIBD) $(BIND)
.PHONY: all clean
