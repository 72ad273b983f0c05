/* rb200_mex_common.h -- shared plumbing of the MEX gateways (one translation unit per gateway).
 *
 * Each gateway keeps the reference function's name and argument/return layout
 *  and is a thin marshalling layer over the C ABI of libradar_b200.so:
 *   - inputs are borrowed, read-only, MATLAB double (split real/imag, column-major);
 *   - outputs are allocated with mxCreateDoubleMatrix and owned by the interpreter;
 *   - one lazily acquired reference to the library's shared context (device RB200_DEVICE or 0) lives across
 *     calls and is released in mexAtExit; all gateways of a session share that one context;
 *   - errors the M-code would raise become mexErrMsgIdAndTxt("radar_b200:<stage>:<what>", ...);
 *     because that call longjmps, every C++ temporary is released before it is reached
 *     (the RB_FAIL macro is only used at points where no non-trivial destructor is pending).
 * Compiles against MATLAB's / Octave's mex.h or against mex/shim/mex.h (tests).
 */
#ifndef RB200_MEX_COMMON_H
#define RB200_MEX_COMMON_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mex.h"
#include "../include/radar_b200.h"
#include "rb200_waveform_literals.h"

/* The context is the library's process-wide shared one (rb200_shared_context_acquire): every gateway binary loaded into
 * the session uses the same CUDA context, device scratch and plan cache; this translation unit only holds a reference. */
static rb200_ctx* g_ctx = NULL;
static int g_dev = 0;

static void rb_shutdown(void) {
    if (g_ctx) rb200_shared_context_release(g_dev);
    g_ctx = NULL;
}

static rb200_ctx* rb_context(void) {
    if (!g_ctx) {
        const char* dev = getenv("RB200_DEVICE");      /* once per gateway, at its first call */
        g_dev = dev ? atoi(dev) : 0;
        if (rb200_shared_context_acquire(&g_ctx, g_dev) != RB200_OK) {
            g_ctx = NULL;
            mexErrMsgIdAndTxt("radar_b200:init:noDevice", "%s", rb200_last_error(NULL));
        }
        mexAtExit(rb_shutdown);
        mexLock();
    }
    return g_ctx;
}

/* translate a library status into the error MATLAB would have raised */
static void rb_check(int st, const char* stage) {
    char id[96];
    if (st == RB200_OK) return;
    const char* what = st == RB200_ERR_INDEX ? "indexOutOfRange" : st == RB200_ERR_DIM_MISMATCH ? "dimensionMismatch"
                     : st == RB200_ERR_ARG ? "badArgument" : st == RB200_ERR_UNSUPPORTED ? "unsupported"
                     : st == RB200_ERR_NO_WAVEFORM ? "noWaveform" : st == RB200_ERR_CUDA ? "cuda" : "error";
    snprintf(id, sizeof id, "radar_b200:%s:%s", stage, what);
    mexErrMsgIdAndTxt(id, "%s", rb200_last_error(g_ctx));
}

static void rb_require(int cond, const char* id, const char* msg) {
    if (!cond) mexErrMsgIdAndTxt(id, "%s", msg);
}

static void rb_require_real_or_complex_double(const mxArray* a, const char* id) {
    rb_require(a && mxIsDouble(a), id, "argument must be a double array");
}

static double rb_scalar(const mxArray* a, const char* id) {
    rb_require(a && mxIsDouble(a) && mxGetNumberOfElements(a) >= 1, id, "expected a numeric scalar");
    return mxGetScalar(a);
}

static int rb_nonzero(const mxArray* a) {
    size_t i, n;
    if (!a || !mxIsDouble(a)) return 0;
    n = mxGetNumberOfElements(a);
    for (i = 0; i < n; ++i)
        if (mxGetPr(a)[i] != 0.0) return 1;
    return 0;
}

static void rb_no_plots(const mxArray* flag) {
    if (rb_nonzero(flag))
        mexWarnMsgIdAndTxt("radar_b200:plot:notReproduced", "show_PC/show_FFT/graph plotting is not reproduced by the MEX gateway");
}

/* MATLAB a:d:b as MathWorks' published COLONOP replication builds it: rb_colon_count = number of elements, rb_colon_fill writes
 * them symmetrically from both ends (a + k d up to the middle, c - k d down from the right end, the exact mid-point for even
 * n), so that the ideal LFM pulses of MTD/fun_MTD_produce.m:62-69 carry MATLAB's own last bits. */
static long rb_colon_n(double a, double d, double b, double* c_out) {
    if (!(a == a) || !(d == d) || !(b == b) || d == 0 || (a < b && d < 0) || (b < a && d > 0)) return -1;
    const double eps = 2.220446049250313e-16;
    const double tol = 2.0 * eps * (fabs(a) > fabs(b) ? fabs(a) : fabs(b));
    const double sig = d > 0 ? 1.0 : -1.0;
    long n;
    if (a == floor(a) && d == 1) n = (long)(floor(b) - a);
    else if (a == floor(a) && d == floor(d)) { const double q = floor(a / d), r = a - q * d; n = (long)(floor((b - r) / d) - q); }
    else { n = (long)floor((b - a) / d + 0.5); if (sig * (a + (double)n * d - b) > tol) n -= 1; }
    if (n < 0) return -1;
    double c = a + (double)n * d;
    if (sig * (c - b) > -tol) c = b;
    if (c_out) *c_out = c;
    return n;
}
static size_t rb_colon_count(double a, double d, double b) {
    const long n = rb_colon_n(a, d, b, NULL);
    return n < 0 ? 0 : (size_t)n + 1;
}
static void rb_colon_fill(double a, double d, double b, double* out) {
    double c;
    const long n = rb_colon_n(a, d, b, &c);
    if (n < 0) return;
    for (long k = 0; k <= n / 2; ++k) {
        out[k] = a + (double)k * d;
        out[n - k] = c - (double)k * d;
    }
    if (n % 2 == 0) out[n / 2] = (a + c) / 2;
}

/* ---- waveform plans ---------------------------------------------------------------------------- */
typedef struct {
    const double *p2re, *p2im, *p3re, *p3im;
    int n2, n3;
} rb_pulses;

/* FNV-1a over the plan key and the pulses: the tag stored in the (shared) context names the resident plan */
static uint64_t rb_plan_tag(const char* key, const rb_pulses* p) {
    uint64_t h = 1469598103934665603ull;
    const double zero = 0.0;
    size_t i;
    int k;
#define RB_MIX(ptr, n) do { const unsigned char* b__ = (const unsigned char*)(ptr); size_t j__; \
        for (j__ = 0; j__ < (size_t)(n); ++j__) { h ^= b__[j__]; h *= 1099511628211ull; } } while (0)
    RB_MIX(key, strlen(key));
    for (k = 0; k < 2; ++k) {
        const double* re = k ? p->p3re : p->p2re;
        const double* im = k ? p->p3im : p->p2im;
        const int n = k ? p->n3 : p->n2;
        RB_MIX(&n, sizeof n);
        for (i = 0; i < (size_t)n; ++i) { RB_MIX(re + i, sizeof(double)); RB_MIX(im ? im + i : &zero, sizeof(double)); }
    }
#undef RB_MIX
    return h ? h : 1;
}

static int rb_same_plan(const char* key, const rb_pulses* p) {
    return rb200_get_plan_tag(rb_context()) == rb_plan_tag(key, p);
}

static void rb_remember_plan(const char* key, const rb_pulses* p) {
    rb200_set_plan_tag(rb_context(), rb_plan_tag(key, p));
}

static void rb_fir_taps(double* b) {   /* filter_coef / max(filter_coef), MP/fun_lss_pulse_compression.m:21-22 */
    int i;
    double mx = RB200_FILTER_COEF[0];
    for (i = 1; i < 35; ++i) if (RB200_FILTER_COEF[i] > mx) mx = RB200_FILTER_COEF[i];
    for (i = 0; i < 35; ++i) b[i] = RB200_FILTER_COEF[i] / mx;
}

/* 5-arg rule (MP/fun_lss_pulse_compression.m:6-8,31,36-37): segments 82/242/rest, offsets 75/160 */
static void rb_plan_mp(int n, const rb_pulses* p) {
    char key[64];
    double b[35];
    rb200_segment s[3];
    rb200_ctx* ctx = rb_context();
    rb_require(n >= 324, "radar_b200:pc:indexOutOfRange", "fun_lss_pulse_compression: Index in position 2 exceeds array bounds (PRT shorter than 324)");
    rb_require(p->n2 == 75, "radar_b200:pc:dimensionMismatch", "fun_lss_pulse_compression: pulse2 must have 75 samples (signal_PC_02(75:end))");
    rb_require(p->n3 == 160, "radar_b200:pc:dimensionMismatch", "fun_lss_pulse_compression: pulse3 must have 160 samples (signal_PC_03(160:end))");
    snprintf(key, sizeof key, "mp:%d", n);
    if (rb_same_plan(key, p)) return;
    rb_fir_taps(b);
    memset(s, 0, sizeof s);
    s[0].in_start = 0; s[0].in_len = 82; s[0].out_start = 0; s[0].out_len = 82;
    s[0].kind = RB200_SEG_FIR; s[0].align = RB200_ALIGN_DELAYED; s[0].n_taps = 35; s[0].taps_re = b; s[0].scale = 1.0 / 1.2;
    s[1].in_start = 82; s[1].in_len = 242; s[1].out_start = 82; s[1].out_len = 242;
    s[1].kind = RB200_SEG_MF; s[1].align = RB200_ALIGN_LEADING_EDGE; s[1].n_taps = p->n2; s[1].taps_re = p->p2re; s[1].taps_im = p->p2im; s[1].scale = 1.0;
    s[2].in_start = 324; s[2].in_len = n - 324; s[2].out_start = 324; s[2].out_len = n - 324;
    s[2].kind = RB200_SEG_MF; s[2].align = RB200_ALIGN_LEADING_EDGE; s[2].n_taps = p->n3; s[2].taps_re = p->p3re; s[2].taps_im = p->p3im; s[2].scale = 1.0;
    rb_check(rb200_set_waveform(ctx, s, 3), "pc");
    rb_remember_plan(key, p);
}

/* 9-arg rule (MTD/fun_lss_pulse_compression.m:23-25,47-51,58-65) */
static void rb_plan_mtd(int n, const rb_pulses* p, int p1, int p2, int p3) {
    char key[64];
    double b[35];
    rb200_segment s[3];
    rb200_ctx* ctx = rb_context();
    rb_require(p1 >= 0 && p2 >= 0 && p3 >= 0 && n >= p1 + p2, "radar_b200:pc:indexOutOfRange", "fun_lss_pulse_compression: Index in position 2 exceeds array bounds");
    rb_require(p3 <= n - p1 - p2, "radar_b200:pc:indexOutOfRange", "fun_lss_pulse_compression: Index exceeds the number of array elements (point_prt3 too large)");
    snprintf(key, sizeof key, "mtd:%d:%d:%d:%d", n, p1, p2, p3);
    if (rb_same_plan(key, p)) return;
    rb_fir_taps(b);
    memset(s, 0, sizeof s);
    s[0].in_start = 0; s[0].in_len = p1; s[0].out_start = 0; s[0].out_len = p1;
    s[0].kind = RB200_SEG_FIR; s[0].align = RB200_ALIGN_GRPDELAY; s[0].n_taps = 35; s[0].taps_re = b; s[0].scale = 1.0 / 1.2;
    s[1].in_start = p1; s[1].in_len = p2; s[1].out_start = p1; s[1].out_len = p2;
    s[1].kind = RB200_SEG_MF; s[1].align = RB200_ALIGN_LEADING_EDGE; s[1].n_taps = p->n2; s[1].taps_re = p->p2re; s[1].taps_im = p->p2im; s[1].scale = 1.0;
    s[2].in_start = p1 + p2; s[2].in_len = n - p1 - p2; s[2].out_start = p1 + p2; s[2].out_len = p3;
    s[2].kind = RB200_SEG_MF; s[2].align = RB200_ALIGN_LEADING_EDGE; s[2].n_taps = p->n3; s[2].taps_re = p->p3re; s[2].taps_im = p->p3im; s[2].scale = 1.0;
    rb_check(rb200_set_waveform(ctx, s, 3), "pc");
    rb_remember_plan(key, p);
}

#endif /* RB200_MEX_COMMON_H */
