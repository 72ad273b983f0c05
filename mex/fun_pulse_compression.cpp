/* signal_PC = fun_pulse_compression(s0, s_echo)     -- MEX gateway
 * Replaces MatlabProcess_xuzerui/fun_pulse_compression.m:1 (identical copy MTD/fun_pulse_compression.m:10).
 * s0: 1xL reference pulse, s_echo: 1xM echo (real or complex double) -> 1x(L+M-1) complex double. */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 2, "radar_b200:pc:nargin", "fun_pulse_compression: expected 2 inputs (s0, s_echo)");
    rb_require(nlhs <= 1, "radar_b200:pc:nargout", "fun_pulse_compression: one output");
    rb_require_real_or_complex_double(prhs[0], "radar_b200:pc:type");
    rb_require_real_or_complex_double(prhs[1], "radar_b200:pc:type");
    /* s0(1,end:-1:1) and size(s_echo,2): both are treated as row vectors */
    const int L = (int)mxGetN(prhs[0]);
    const int M = (int)mxGetN(prhs[1]);
    rb_require(mxGetM(prhs[0]) >= 1 && L >= 1, "radar_b200:pc:indexOutOfRange", "fun_pulse_compression: s0 is empty");
    const int n = L + M - 1;
    plhs[0] = mxCreateDoubleMatrix(1, n > 0 ? n : 0, mxCOMPLEX);
    if (n <= 0) return;
    rb_check(rb200_pulse_compression_z(rb_context(), mxGetPr(prhs[0]), mxGetPi(prhs[0]), L, mxGetPr(prhs[1]), mxGetPi(prhs[1]), M,
                                       mxGetPr(plhs[0]), mxGetPi(plhs[0])), "pc");
}
