/* MTD = fun_0v_pressing(MTD)     -- MEX gateway
 * Replaces MatlabProcess_xuzerui/fun_0v_pressing.m:2 and MTD/fun_0v_pressing.m:13 (divisor 150).
 * Built a second time with -DRB200_ZERO_V_DIV=20 for MatlabProcess_xuzerui/CFAR_WangCai/fun_0v_pressing.m:2. */
#include "rb200_mex_common.h"
#ifndef RB200_ZERO_V_DIV
#define RB200_ZERO_V_DIV 150
#endif

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 1, "radar_b200:zerov:nargin", "fun_0v_pressing: expected 1 input");
    rb_require(nlhs <= 1, "radar_b200:zerov:nargout", "fun_0v_pressing: one output");
    rb_require(prhs[0] && mxIsDouble(prhs[0]) && !mxIsComplex(prhs[0]), "radar_b200:zerov:type", "fun_0v_pressing: MTD must be a real double matrix");
    const int P = (int)mxGetM(prhs[0]), R = (int)mxGetN(prhs[0]);
    plhs[0] = mxCreateDoubleMatrix(P, R, mxREAL);
    rb_require(P >= 1, "radar_b200:zerov:indexOutOfRange", "fun_0v_pressing: Index in position 1 is invalid (empty MTD)");
    rb_check(rb200_zero_v_pressing_d(rb_context(), mxGetPr(prhs[0]), P, R, RB200_ZERO_V_DIV, mxGetPr(plhs[0])), "zerov");
}
