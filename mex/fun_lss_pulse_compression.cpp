/* s_PC_0 = fun_lss_pulse_compression(echo, show_PC, pulse1, pulse2, pulse3)                       (5-arg)
 * s_PC_0 = fun_lss_pulse_compression(echo, params, show_PC, pulse1, pulse2, pulse3, p1, p2, p3)   (9-arg)
 * Replaces MatlabProcess_xuzerui/fun_lss_pulse_compression.m:3 and MTD/fun_lss_pulse_compression.m:17;
 * the gateway dispatches on nargin because the two generations differ in segment sizes and in the
 * alignment of the short-pulse FIR output. */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 5 || nrhs == 9, "radar_b200:pc:nargin", "fun_lss_pulse_compression: expected 5 or 9 inputs");
    rb_require(nlhs <= 1, "radar_b200:pc:nargout", "fun_lss_pulse_compression: one output");
    const mxArray* echo = prhs[0];
    rb_require_real_or_complex_double(echo, "radar_b200:pc:type");
    const int P = (int)mxGetM(echo), R = (int)mxGetN(echo);
    const int o = nrhs == 5 ? 1 : 2;                 /* index of show_PC */
    rb_no_plots(prhs[o]);
    const mxArray *pulse2 = prhs[o + 2], *pulse3 = prhs[o + 3];
    rb_require_real_or_complex_double(pulse2, "radar_b200:pc:type");
    rb_require_real_or_complex_double(pulse3, "radar_b200:pc:type");
    rb_pulses pl;
    pl.p2re = mxGetPr(pulse2); pl.p2im = mxGetPi(pulse2); pl.n2 = (int)mxGetNumberOfElements(pulse2);
    pl.p3re = mxGetPr(pulse3); pl.p3im = mxGetPi(pulse3); pl.n3 = (int)mxGetNumberOfElements(pulse3);
    if (nrhs == 5) rb_plan_mp(R, &pl);
    else rb_plan_mtd(R, &pl, (int)rb_scalar(prhs[6], "radar_b200:pc:type"), (int)rb_scalar(prhs[7], "radar_b200:pc:type"),
                     (int)rb_scalar(prhs[8], "radar_b200:pc:type"));
    plhs[0] = mxCreateDoubleMatrix(P, R, mxCOMPLEX);
    if (P == 0 || R == 0) return;
    rb_check(rb200_lss_pulse_compression_z(rb_context(), mxGetPr(echo), mxGetPi(echo), P, R, mxGetPr(plhs[0]), mxGetPi(plhs[0])), "pc");
}
