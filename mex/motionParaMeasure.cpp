/* [rEstSeries, vEstSeries, eleAngleEstSeries] = motionParaMeasure(echo_MTD_sum_short, echo_MTD_diff_short,
 *      cfarResultFlag_Matrix_short, extraDots, rScale_short, deltaR, rInterpTimes, vScale, deltaV, vInterpTimes,
 *      kValues, beamPosNum, beamAngleStep, freInd, eleAngleComp, eleAngleSysErr, MTD_0_num)         -- MEX gateway
 * Replaces MatlabProcess_xuzerui/CFAR_WangCai/motionParaMeasure.m:1.
 * Outputs are column vectors with one entry per flagged cell in find() order (empty when nothing is flagged). */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 17, "radar_b200:measure:nargin", "motionParaMeasure: expected 17 inputs");
    rb_require(nlhs <= 3, "radar_b200:measure:nargout", "motionParaMeasure: at most three outputs");
    for (int i = 0; i < 3; ++i)
        rb_require(prhs[i] && mxIsDouble(prhs[i]) && !mxIsComplex(prhs[i]), "radar_b200:measure:type", "motionParaMeasure: matrices must be real double");
    const int V = (int)mxGetM(prhs[2]), R = (int)mxGetN(prhs[2]);
    rb_require((int)mxGetM(prhs[0]) >= V && (int)mxGetN(prhs[0]) >= R && (int)mxGetM(prhs[1]) >= V && (int)mxGetN(prhs[1]) >= R &&
               (int)mxGetM(prhs[0]) == (int)mxGetM(prhs[1]), "radar_b200:measure:indexOutOfRange",
               "motionParaMeasure: Index exceeds array bounds (sum/diff smaller than the flag matrix)");
    rb_require((int)mxGetM(prhs[0]) == V, "radar_b200:measure:unsupported", "motionParaMeasure: sum/diff must have as many rows as the flag matrix");
    rb_require((int)mxGetNumberOfElements(prhs[4]) >= R && (int)mxGetNumberOfElements(prhs[7]) >= V, "radar_b200:measure:indexOutOfRange",
               "motionParaMeasure: rScale / vScale shorter than the matrix");
    size_t cap = 0;
    const double* f = mxGetPr(prhs[2]);
    for (size_t i = 0; i < (size_t)V * R; ++i) cap += f[i] != 0.0;
    mxArray* o[3];
    for (int i = 0; i < 3; ++i) o[i] = mxCreateDoubleMatrix(cap, cap ? 1 : 0, mxREAL);
    int n = 0, st = RB200_OK;
    if (V >= 1 && R >= 1)
        st = rb200_motion_para_measure_d(rb_context(), mxGetPr(prhs[0]), mxGetPr(prhs[1]), f, V, R, (int)rb_scalar(prhs[3], "radar_b200:measure:type"),
                                         mxGetPr(prhs[4]), rb_scalar(prhs[5], "radar_b200:measure:type"), (int)rb_scalar(prhs[6], "radar_b200:measure:type"),
                                         mxGetPr(prhs[7]), rb_scalar(prhs[8], "radar_b200:measure:type"), (int)rb_scalar(prhs[9], "radar_b200:measure:type"),
                                         mxGetPr(prhs[10]), (int)mxGetM(prhs[10]), (int)mxGetN(prhs[10]), rb_scalar(prhs[11], "radar_b200:measure:type"),
                                         rb_scalar(prhs[12], "radar_b200:measure:type"), (int)rb_scalar(prhs[13], "radar_b200:measure:type"),
                                         rb_scalar(prhs[14], "radar_b200:measure:type"), rb_scalar(prhs[15], "radar_b200:measure:type"),
                                         (int)rb_scalar(prhs[16], "radar_b200:measure:type"), mxGetPr(o[0]), mxGetPr(o[1]), mxGetPr(o[2]), (int)cap, &n);
    plhs[0] = o[0];
    if (nlhs >= 2) plhs[1] = o[1]; else mxDestroyArray(o[1]);
    if (nlhs >= 3) plhs[2] = o[2]; else mxDestroyArray(o[2]);
    rb_check(st, "measure");
}
