# Builds libbbme.so (sm_100a only) in-tree, the CPU oracle, and (when /root/reference exists) oracle/_ref.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := blockbasedmotionestimation_b200
CSRC := $(PKG)/csrc
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function \
           --fmad=false -Iinclude
LIB := $(PKG)/libbbme.so
OBJS := $(CSRC)/kernels.o $(CSRC)/regularize.o $(CSRC)/search_tma.o $(CSRC)/capi.o $(CSRC)/flo.o $(CSRC)/hostpool.o $(CSRC)/multi.o

all: $(LIB) oracle tools/color_flow

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/kernels.h $(CSRC)/hostpool.h include/bbme.h
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@

$(CSRC)/flo.o: $(CSRC)/flo.cpp include/bbme.h
	g++ -O2 -ffp-contract=off -fPIC -std=c++17 -Wall -Iinclude -c $< -o $@

$(CSRC)/hostpool.o: $(CSRC)/hostpool.cpp $(CSRC)/hostpool.h
	g++ -O3 -fPIC -std=c++17 -Wall -c $< -o $@

$(CSRC)/multi.o: $(CSRC)/multi.cpp include/bbme.h
	g++ -O2 -fPIC -std=c++17 -Wall -Iinclude -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(OBJS) -cudart static -lpthread

# color_flow-compatible command-line tool (middlebury/flow-code/color_flow.cpp:68-98); host-only
tools/color_flow: tools/color_flow.cpp include/bbme.h $(LIB)
	g++ -O2 -std=c++17 -Wall -Iinclude -o $@ tools/color_flow.cpp -L$(PKG) -lbbme -Wl,-rpath,'$$ORIGIN/../$(PKG)'

oracle:
	$(MAKE) -C oracle

micro: bench_micro/int_peak bench_micro/shift_pipe
bench_micro/%: bench_micro/%.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o $@ $<

clean:
	rm -f $(OBJS) $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean micro
