/* PeakDetectionoutput = Function_CFAR1D_sub_fixCells(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod, rowCellsFix, colCellsFix)
 * Replaces MatlabProcess_xuzerui/CFAR_WangCai/Function_CFAR1D_sub_fixCells.m:1 (1-based row / column lists). */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 7, "radar_b200:cfar1d:nargin", "Function_CFAR1D_sub_fixCells: expected 7 inputs");
    rb_require(nlhs <= 1, "radar_b200:cfar1d:nargout", "Function_CFAR1D_sub_fixCells: one output");
    rb_require(prhs[0] && mxIsDouble(prhs[0]) && !mxIsComplex(prhs[0]), "radar_b200:cfar1d:type", "Function_CFAR1D_sub_fixCells: datamatrix must be real double");
    const int rows = (int)mxGetM(prhs[0]), cols = (int)mxGetN(prhs[0]);
    const int ref = (int)rb_scalar(prhs[1], "radar_b200:cfar1d:type"), guard = (int)rb_scalar(prhs[2], "radar_b200:cfar1d:type");
    const double T = rb_scalar(prhs[3], "radar_b200:cfar1d:type");
    const int method = (int)rb_scalar(prhs[4], "radar_b200:cfar1d:type");
    const size_t nr = mxGetNumberOfElements(prhs[5]), nc = mxGetNumberOfElements(prhs[6]);
    int32_t* idx = (int32_t*)malloc((nr + nc + 1) * sizeof(int32_t));
    for (size_t i = 0; i < nr; ++i) idx[i] = (int32_t)mxGetPr(prhs[5])[i];
    for (size_t i = 0; i < nc; ++i) idx[nr + i] = (int32_t)mxGetPr(prhs[6])[i];
    plhs[0] = mxCreateDoubleMatrix(rows, cols, mxREAL);
    int st = RB200_OK;
    if (rows > 0 && cols > 0)
        st = rb200_cfar1d_fix_d(rb_context(), mxGetPr(prhs[0]), rows, cols, ref, guard, T, method != 0, idx, (int)nr, idx + nr, (int)nc, mxGetPr(plhs[0]));
    free(idx);
    rb_check(st, "cfar1d");
}
