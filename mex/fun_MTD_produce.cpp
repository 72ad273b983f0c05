/* MTD_Signal = fun_MTD_produce(echo)             MatlabProcess_xuzerui/fun_MTD_produce.m:3  (literal pulses, 5-arg PC)
 * MTD_Signal = fun_MTD_produce(echo, params)     MTD/fun_MTD_produce.m:12                   (ideal LFM pulses, 9-arg PC)
 * Pulse compression -> fun_Process_MTD (Kaiser beta 8) -> fun_0v_pressing (divisor 150), one library call. */
#include "rb200_mex_common.h"

static const mxArray* field(const mxArray* s, const char* name) {
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("radar_b200:mtdproduce:missingField", "fun_MTD_produce: params.%s is missing", name);
    return f;
}

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 1 || nrhs == 2, "radar_b200:mtdproduce:nargin", "fun_MTD_produce: expected (echo) or (echo, params)");
    rb_require(nlhs <= 1, "radar_b200:mtdproduce:nargout", "fun_MTD_produce: one output");
    const mxArray* echo = prhs[0];
    rb_require_real_or_complex_double(echo, "radar_b200:mtdproduce:type");
    const int P = (int)mxGetM(echo), R = (int)mxGetN(echo);
    rb_require(P >= 1 && R >= 1, "radar_b200:mtdproduce:badArgument", "fun_MTD_produce: echo is empty");
    rb_pulses pl;
    double *q2re = NULL, *q2im = NULL, *q3re = NULL, *q3im = NULL;
    if (nrhs == 1) {
        pl.p2re = RB200_PULSE2_REAL; pl.p2im = RB200_PULSE2_IMAG; pl.n2 = 75;       /* fun_MTD_produce.m:54-56 */
        pl.p3re = RB200_PULSE3_REAL; pl.p3im = RB200_PULSE3_IMAG; pl.n3 = 160;      /* :58-60 */
        rb_plan_mp(R, &pl);
    } else {
        const mxArray* prm = prhs[1];
        rb_require(mxIsStruct(prm), "radar_b200:mtdproduce:type", "fun_MTD_produce: params must be a struct");
        const mxArray* dbg = mxGetField(prm, 0, "debug");
        if (dbg && mxIsStruct(dbg)) {
            rb_no_plots(mxGetField(dbg, 0, "show_PC"));
            rb_no_plots(mxGetField(dbg, 0, "show_FFT"));
            rb_no_plots(mxGetField(dbg, 0, "graph"));
        }
        const double fs = rb_scalar(field(prm, "fs"), "radar_b200:mtdproduce:type");
        const double B = rb_scalar(field(prm, "B"), "radar_b200:mtdproduce:type");
        const mxArray* tao = field(prm, "tao");
        const mxArray* pp = field(prm, "point_prt");
        rb_require(mxGetNumberOfElements(tao) >= 3 && mxGetNumberOfElements(pp) >= 4, "radar_b200:mtdproduce:indexOutOfRange",
                   "fun_MTD_produce: params.tao needs 3 and params.point_prt 4 elements");
        const double ts = 1.0 / fs;
        const double tao2 = mxGetPr(tao)[1], tao3 = mxGetPr(tao)[2];
        const double K2 = -B / tao2, K3 = B / tao3;                                  /* MTD/fun_MTD_produce.m:50-51 */
        const size_t n2 = rb_colon_count(-tao2 / 2, ts, tao2 / 2 - ts), n3 = rb_colon_count(-tao3 / 2, ts, tao3 / 2 - ts);
        q2re = (double*)malloc((n2 + 1) * sizeof(double)); q2im = (double*)malloc((n2 + 1) * sizeof(double));
        q3re = (double*)malloc((n3 + 1) * sizeof(double)); q3im = (double*)malloc((n3 + 1) * sizeof(double));
        rb_colon_fill(-tao2 / 2, ts, tao2 / 2 - ts, q2re);                           /* t2 = -tao2/2:ts:tao2/2-ts, :62 */
        rb_colon_fill(-tao3 / 2, ts, tao3 / 2 - ts, q3re);                           /* :63 */
        for (size_t i = 0; i < n2; ++i) {                                            /* :68 */
            const double t = q2re[i], ph = 2.0 * M_PI * (0.5 * K2 * (t * t));
            q2re[i] = cos(ph); q2im[i] = sin(ph);
        }
        for (size_t i = 0; i < n3; ++i) {                                            /* :69 */
            const double t = q3re[i], ph = 2.0 * M_PI * (0.5 * K3 * (t * t));
            q3re[i] = cos(ph); q3im[i] = sin(ph);
        }
        pl.p2re = q2re; pl.p2im = q2im; pl.n2 = (int)n2;
        pl.p3re = q3re; pl.p3im = q3im; pl.n3 = (int)n3;
        const int p1 = (int)mxGetPr(pp)[1], p2 = (int)mxGetPr(pp)[2], p3 = (int)mxGetPr(pp)[3];
        /* argument errors must not leak the malloc'ed pulses: validate before planning */
        const int bad = (p1 < 0 || p2 < 0 || p3 < 0 || R < p1 + p2 || p3 > R - p1 - p2 || n2 < 1 || n3 < 1);
        if (bad) { free(q2re); free(q2im); free(q3re); free(q3im); }
        rb_require(!bad, "radar_b200:pc:indexOutOfRange", "fun_lss_pulse_compression: Index exceeds array bounds (point_prt does not fit the PRT)");
        rb_plan_mtd(R, &pl, p1, p2, p3);
        free(q2re); free(q2im); free(q3re); free(q3im);
    }
    plhs[0] = mxCreateDoubleMatrix(P, R, mxREAL);
    rb_check(rb200_mtd_produce_z(rb_context(), mxGetPr(echo), mxGetPi(echo), P, R, 8.0, 150, mxGetPr(plhs[0])), "mtdproduce");
}
