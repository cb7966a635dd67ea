# Builds the product library (CUDA, sm_100a) and the CPU oracle (test infrastructure).
# nvcc cross-compiles without a GPU.  Outputs are git-ignored but travel to the GPU box.
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
# --fmad=false: the bit-exact float paths (LBD, fastAtan2, resize) mirror CPU code compiled
# without FMA contraction; FMA is requested explicitly (__fma_rn/__fmaf_rn) where wanted.
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC -Iinclude -Ipl-slam_b200/csrc \
             -Xptxas -warn-spills -Wno-deprecated-gpu-targets
LIBDIR    := pl-slam_b200/lib
OBJDIR    := build/obj
SRCS      := $(wildcard pl-slam_b200/csrc/*.cu)
OBJS      := $(patsubst pl-slam_b200/csrc/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := $(LIBDIR)/libplslam_b200.so

ORACLE_SRCS := $(wildcard oracle/*.c)
ORACLE_LIB  := oracle/_build/liboracle.so

DEMO      := $(LIBDIR)/vo_demo

all: $(LIB) oracle $(DEMO)

$(OBJDIR)/%.o: pl-slam_b200/csrc/%.cu $(wildcard pl-slam_b200/csrc/*.h) $(wildcard pl-slam_b200/csrc/*.cuh) include/plslam_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS)

# C++ host shim demo (the reference's VO loop on the shim classes)
$(DEMO): pl-slam_b200/cpp/vo_demo.cpp pl-slam_b200/cpp/stvo_shim.h include/plslam_b200.h $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -Ipl-slam_b200/cpp -o $@ pl-slam_b200/cpp/vo_demo.cpp -L$(LIBDIR) -lplslam_b200 -Wl,-rpath,'$$ORIGIN'

# CPU oracle: plain C, no FMA contraction (mirrors the reference's non-FMA x86-64 build).
oracle: $(ORACLE_LIB)
$(ORACLE_LIB): $(ORACLE_SRCS) $(wildcard oracle/*.h)
	@mkdir -p oracle/_build
	$(CC) -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -std=c11 -Wall -o $@ $(ORACLE_SRCS) -lm -lpthread

clean:
	rm -rf build $(LIBDIR)/*.so oracle/_build

.PHONY: all oracle clean
