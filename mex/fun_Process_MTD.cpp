/* MTD_Signal = fun_Process_MTD(ProSignal, Len_PRT, Num_PRTperFrame)     -- MEX gateway
 * Replaces MatlabProcess_xuzerui/fun_Process_MTD.m:3 (identical copy MTD/fun_Process_MTD.m:13):
 * kaiser(Num_PRTperFrame, 8) window, slow-time FFT, fftshift, abs. */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 3, "radar_b200:mtd:nargin", "fun_Process_MTD: expected 3 inputs (ProSignal, Len_PRT, Num_PRTperFrame)");
    rb_require(nlhs <= 1, "radar_b200:mtd:nargout", "fun_Process_MTD: one output");
    rb_require_real_or_complex_double(prhs[0], "radar_b200:mtd:type");
    const int rows = (int)mxGetM(prhs[0]), cols = (int)mxGetN(prhs[0]);
    const int len_prt = (int)rb_scalar(prhs[1], "radar_b200:mtd:type");
    const int num_prt = (int)rb_scalar(prhs[2], "radar_b200:mtd:type");
    rb_require(num_prt >= 1 && len_prt >= 0, "radar_b200:mtd:badArgument", "fun_Process_MTD: sizes must be positive");
    plhs[0] = mxCreateDoubleMatrix(num_prt, len_prt, mxREAL);
    rb_check(rb200_process_mtd_z(rb_context(), mxGetPr(prhs[0]), mxGetPi(prhs[0]), rows, cols, len_prt, num_prt, 8.0, mxGetPr(plhs[0])), "mtd");
}
