/* PeakDetectionoutput = Function_CFAR1D_sub(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod)
 * Replaces MatlabProcess_xuzerui/CFAR_WangCai/Function_CFAR1D_sub.m:1 (detection along the column index). */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 5, "radar_b200:cfar1d:nargin", "Function_CFAR1D_sub: expected 5 inputs");
    rb_require(nlhs <= 1, "radar_b200:cfar1d:nargout", "Function_CFAR1D_sub: one output");
    rb_require(prhs[0] && mxIsDouble(prhs[0]) && !mxIsComplex(prhs[0]), "radar_b200:cfar1d:type", "Function_CFAR1D_sub: datamatrix must be real double");
    const int rows = (int)mxGetM(prhs[0]), cols = (int)mxGetN(prhs[0]);
    const int ref = (int)rb_scalar(prhs[1], "radar_b200:cfar1d:type"), guard = (int)rb_scalar(prhs[2], "radar_b200:cfar1d:type");
    const double T = rb_scalar(prhs[3], "radar_b200:cfar1d:type");
    const int method = (int)rb_scalar(prhs[4], "radar_b200:cfar1d:type");
    plhs[0] = mxCreateDoubleMatrix(rows, cols, mxREAL);
    if (rows == 0 || cols == 0) return;
    rb_check(rb200_cfar1d_sub_d(rb_context(), mxGetPr(prhs[0]), rows, cols, ref, guard, T, method != 0, mxGetPr(plhs[0])), "cfar1d");
}
