# Builds libradar_b200.so (sm_100a only) and the MEX shim harness.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= nvcc
CXX       ?= g++
PKG       := radar_signal_process_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libradar_b200.so
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas --expt-relaxed-constexpr
CU_SRCS   := $(CSRC)/api.cu $(CSRC)/pc_kernels.cu $(CSRC)/mtd_kernels.cu $(CSRC)/mtd64_kernel.cu $(CSRC)/cfar_kernels.cu $(CSRC)/layout_kernels.cu
CU_OBJS   := $(CU_SRCS:.cu=.o)
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/radar_b200.h

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXTRA_NVFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS)
	$(NVCC) -shared -o $@ $(CU_OBJS) -gencode arch=compute_100a,code=sm_100a -cudart shared

clean:
	rm -f $(CU_OBJS) $(LIB)

.PHONY: all clean
