/* [cfarResultFlag_Matrix, cfarResultFlag_MatrixV] = executeCFAR(echo_MTD, refCells_R, saveCells_R, T_CFAR_R, CFARmethod_R,
 *                                                   refCells_V, saveCells_V, T_CFAR_V, CFARmethod_V, MTD_0_num, rCFARDetect_Flag)
 * Replaces MatlabProcess_xuzerui/CFAR_WangCai/executeCFAR.m:1.  Outputs are V x R double 0/1 matrices. */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 11, "radar_b200:cfar:nargin", "executeCFAR: expected 11 inputs");
    rb_require(nlhs <= 2, "radar_b200:cfar:nargout", "executeCFAR: at most two outputs");
    rb_require(prhs[0] && mxIsDouble(prhs[0]) && !mxIsComplex(prhs[0]), "radar_b200:cfar:type", "executeCFAR: MTD must be a real double matrix");
    double a[10];
    for (int i = 0; i < 10; ++i) a[i] = rb_scalar(prhs[1 + i], "radar_b200:cfar:type");
    const int V = (int)mxGetM(prhs[0]), R = (int)mxGetN(prhs[0]);
    plhs[0] = mxCreateDoubleMatrix(V, R, mxREAL);
    mxArray* fv = mxCreateDoubleMatrix(V, R, mxREAL);
    if (nlhs >= 2) plhs[1] = fv;
    int st = RB200_OK;
    if (V >= 1 && R >= 1)
        st = rb200_execute_cfar_d(rb_context(), mxGetPr(prhs[0]), V, R, (int)a[0], (int)a[1], a[2], (int)a[3], (int)a[4], (int)a[5], a[6],
                                  (int)a[7], (int)a[8], a[9] != 0.0, mxGetPr(plhs[0]), mxGetPr(fv));
    if (nlhs < 2) mxDestroyArray(fv);
    rb_check(st, "cfar");
}
