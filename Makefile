# Top-level Makefile for C++ users of the B200 backend (the Python entry points do the same:
# __graft_entry__.build(), fidibench_b200/build.py, drivers/build.py).
#
#   make            libfidib200.so + the four drivers
#   make lib        fidibench_b200/lib/libfidib200.so   (nvcc, sm_100a; cross-compiles without a GPU)
#   make drivers    drivers/bin/{upwindCuda,laplacianCuda,upwindMpiCuda,testStencil2dCuda}
#   make oracle     the test-only CPU oracle (+ oracle/_ref where the reference tree exists)
#   make test       the CPU test suite;   make test-gpu   the parity suite (needs a B200)
NVCC ?= $(firstword $(wildcard /usr/local/cuda/bin/nvcc) nvcc)
CXX  := $(firstword $(wildcard /usr/bin/g++) g++)
ROOT := $(dir $(abspath $(lastword $(MAKEFILE_LIST))))
CSRC := $(ROOT)fidibench_b200/csrc
LIB  := $(ROOT)fidibench_b200/lib/libfidib200.so
OBJD := $(ROOT)fidibench_b200/build/make
SRCS := runtime.cu kernels_generic.cu kernels_tma.cu kernels_fused.cu kernels_lapfused.cu decomp.cu capi.cu
OBJS := $(SRCS:%.cu=$(OBJD)/%.o)
HDRS := $(CSRC)/fdb_internal.h $(CSRC)/tma_ptx.cuh $(ROOT)include/fidib200.h
ARCH := -gencode arch=compute_100a,code=sm_100a
# --fmad=false: the parity contract forbids contracting a*b+c
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo --fmad=false -ccbin $(CXX) -Xcompiler -fPIC,-O2 -I $(ROOT)include -I $(CSRC)
DRIVERS := upwindCuda laplacianCuda upwindMpiCuda testStencil2dCuda
BINS := $(DRIVERS:%=$(ROOT)drivers/bin/%)

.PHONY: all lib drivers oracle test test-gpu clean
all: lib drivers
lib: $(LIB)
drivers: $(BINS)

$(OBJD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -ccbin $(CXX) -shared -o $@ $(OBJS) -lnccl

$(ROOT)drivers/bin/%: $(ROOT)drivers/%.cxx $(ROOT)drivers/cmdline.hpp $(ROOT)drivers/Upwind.hpp $(ROOT)drivers/Filter.hpp $(ROOT)include/fidib200.h $(LIB)
	@mkdir -p $(ROOT)drivers/bin
	$(CXX) -std=c++11 -O2 -Wall -I $(ROOT)include -I $(ROOT)drivers $< -o $@ -L $(dir $(LIB)) -lfidib200 \
	    '-Wl,-rpath,$$ORIGIN/../../fidibench_b200/lib' -Wl,--allow-shlib-undefined

oracle:
	$(MAKE) -C $(ROOT)oracle all

test:
	cd $(ROOT) && python -m pytest tests -q -m "not gpu"
test-gpu:
	cd $(ROOT) && python -m pytest tests -q -m gpu

clean:
	rm -rf $(OBJD)
