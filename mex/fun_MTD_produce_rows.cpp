/* MTD_crop = fun_MTD_produce_rows(echo, row_lo, row_hi)
 * == MTD = fun_MTD_produce(echo); MTD_crop = MTD(row_lo:row_hi, :);      MatlabProcess_xuzerui/main_produce_dataset_win_xzr.m:37-40
 * (the script keeps rows 691:845 of the 1536 Doppler rows).  One library call that runs the slow-time transform first and
 * pulse-compresses only the kept rows (rb200_mtd_produce_rows_z): the two operators are linear and act on different axes. */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 3, "radar_b200:mtdproduce:nargin", "fun_MTD_produce_rows: expected (echo, row_lo, row_hi)");
    rb_require(nlhs <= 1, "radar_b200:mtdproduce:nargout", "fun_MTD_produce_rows: one output");
    const mxArray* echo = prhs[0];
    rb_require_real_or_complex_double(echo, "radar_b200:mtdproduce:type");
    const int P = (int)mxGetM(echo), R = (int)mxGetN(echo);
    rb_require(P >= 1 && R >= 1, "radar_b200:mtdproduce:badArgument", "fun_MTD_produce_rows: echo is empty");
    const int row_lo = (int)rb_scalar(prhs[1], "radar_b200:mtdproduce:type");
    const int row_hi = (int)rb_scalar(prhs[2], "radar_b200:mtdproduce:type");
    rb_require(row_lo >= 1 && row_hi <= P && row_lo <= row_hi, "radar_b200:mtdproduce:indexOutOfRange",
               "fun_MTD_produce_rows: Index in position 1 exceeds array bounds (row crop outside 1..P)");
    rb_pulses pl;
    pl.p2re = RB200_PULSE2_REAL; pl.p2im = RB200_PULSE2_IMAG; pl.n2 = 75;       /* fun_MTD_produce.m:54-56 */
    pl.p3re = RB200_PULSE3_REAL; pl.p3im = RB200_PULSE3_IMAG; pl.n3 = 160;      /* :58-60 */
    rb_plan_mp(R, &pl);
    plhs[0] = mxCreateDoubleMatrix(row_hi - row_lo + 1, R, mxREAL);
    rb_check(rb200_mtd_produce_rows_z(rb_context(), mxGetPr(echo), mxGetPi(echo), P, R, 8.0, 150, row_lo, row_hi, mxGetPr(plhs[0])), "mtdproduce");
}
