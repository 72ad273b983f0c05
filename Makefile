# Builds libradar_b200.so (sm_100a only) and the MEX shim harness.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= nvcc
CXX       ?= g++
PKG       := radar_signal_process_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libradar_b200.so
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas --expt-relaxed-constexpr
CU_SRCS   := $(CSRC)/api.cu $(CSRC)/pc_kernels.cu $(CSRC)/pcw_kernel.cu $(CSRC)/mtd_kernels.cu $(CSRC)/mtd64_kernel.cu $(CSRC)/mtd64_tc_kernel.cu $(CSRC)/chain64_kernel.cu $(CSRC)/onepass_kernel.cu $(CSRC)/dbf_kernel.cu $(CSRC)/measure_kernels.cu $(CSRC)/cfar_kernels.cu $(CSRC)/layout_kernels.cu
CU_OBJS   := $(CU_SRCS:.cu=.o) $(CSRC)/reader.o
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/radar_b200.h

MEXDIR    := mex
MEXBUILD  := mex/build
MEX_NAMES := DMX_frame_process motionParaMeasure fun_MTD_produce fun_MTD_produce_rows fun_lss_pulse_compression fun_pulse_compression fun_Process_MTD fun_0v_pressing executeCFAR Function_CFAR1D_sub Function_CFAR1D_sub_fixCells
MEX_SOS   := $(addprefix $(MEXBUILD)/,$(addsuffix .so,$(MEX_NAMES))) $(MEXBUILD)/fun_0v_pressing_cw.so
SHIM      := $(MEXBUILD)/librbmexshim.so
MEXFLAGS  := -O2 -fPIC -shared -Wall -I$(MEXDIR)/shim -I$(MEXDIR) -Wno-unused-function

all: $(LIB) mex-shim

# MEX gateways compiled against the shim mex.h and linked to the C-ABI library (tests run them without MATLAB/Octave)
mex-shim: $(SHIM) $(MEX_SOS)

$(SHIM): $(MEXDIR)/shim/mex_shim.c $(MEXDIR)/shim/mex.h
	@mkdir -p $(MEXBUILD)
	gcc -O2 -fPIC -shared -Wall -I$(MEXDIR)/shim -o $@ $<

$(MEXBUILD)/%.so: $(MEXDIR)/%.cpp $(MEXDIR)/rb200_mex_common.h $(MEXDIR)/rb200_waveform_literals.h include/radar_b200.h $(LIB) $(SHIM)
	$(CXX) $(MEXFLAGS) -o $@ $< -L$(PKG) -lradar_b200 -L$(MEXBUILD) -lrbmexshim -Wl,-rpath,'$$ORIGIN/../../$(PKG)' -Wl,-rpath,'$$ORIGIN'

$(MEXBUILD)/fun_0v_pressing_cw.so: $(MEXDIR)/fun_0v_pressing.cpp $(MEXDIR)/rb200_mex_common.h include/radar_b200.h $(LIB) $(SHIM)
	$(CXX) $(MEXFLAGS) -DRB200_ZERO_V_DIV=20 -o $@ $< -L$(PKG) -lradar_b200 -L$(MEXBUILD) -lrbmexshim -Wl,-rpath,'$$ORIGIN/../../$(PKG)' -Wl,-rpath,'$$ORIGIN'

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) $(EXTRA_NVFLAGS) -c $< -o $@

$(CSRC)/reader.o: $(CSRC)/reader.cpp include/radar_b200.h
	$(CXX) -O2 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(CU_OBJS)
	$(NVCC) -shared -o $@ $(CU_OBJS) -gencode arch=compute_100a,code=sm_100a -cudart shared

clean:
	rm -f $(CU_OBJS) $(LIB)
	rm -rf $(MEXBUILD)

.PHONY: all clean mex-shim
