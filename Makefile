# Builds the B200-native daligner library (sm_100a only) and the oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas -Xptxas -v
CSRC      := damar_b200/csrc
OBJS      := $(CSRC)/api.o $(CSRC)/scan.o $(CSRC)/radix_sort.o $(CSRC)/kmer_index.o $(CSRC)/seed_merge.o $(CSRC)/tandem.o $(CSRC)/pair_align.o $(CSRC)/trace_pts.o $(CSRC)/las_sink.o $(CSRC)/las_device.o
LIB       := damar_b200/libdamar_b200.so

.PHONY: all lib host oracle clean
all: lib host oracle

host: damar_b200/host/simdb

damar_b200/host/simdb: damar_b200/host/simdb.cpp
	g++ -O2 -o $@ $<

lib: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/wave_core.cuh $(CSRC)/path_fuse_core.cuh include/damar_b200.h
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $@.ptxas.log || (cat $@.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB)
