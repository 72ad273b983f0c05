/* mex_shim.c -- tiny MEX runtime used by the tests to execute the gateways without MATLAB/Octave.
 * Implements the subset declared in mex.h plus a few rbshim_* helpers the Python harness calls
 * through ctypes (struct construction, calling a mexFunction with error capture). */
#include "mex.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_FIELDS 32

struct mxArray_tag {
    int is_struct;
    size_t m, n;
    double* pr;
    double* pi;
    int nfields;
    char* names[MAX_FIELDS];
    mxArray* values[MAX_FIELDS];
};

static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static char g_err_id[256];
static char g_err_msg[1024];
static char g_warn_id[256];
static int g_warn_count = 0;
static void (*g_atexit[16])(void);
static int g_natexit = 0;

double* mxGetPr(const mxArray* a) { return a ? a->pr : NULL; }
double* mxGetPi(const mxArray* a) { return a ? a->pi : NULL; }
size_t mxGetM(const mxArray* a) { return a ? a->m : 0; }
size_t mxGetN(const mxArray* a) { return a ? a->n : 0; }
size_t mxGetNumberOfElements(const mxArray* a) { return a ? a->m * a->n : 0; }
int mxIsComplex(const mxArray* a) { return a && a->pi != NULL; }
int mxIsDouble(const mxArray* a) { return a && !a->is_struct; }
int mxIsStruct(const mxArray* a) { return a && a->is_struct; }
int mxIsEmpty(const mxArray* a) { return !a || a->m * a->n == 0; }
double mxGetScalar(const mxArray* a) { return (a && a->pr && a->m * a->n > 0) ? a->pr[0] : 0.0; }

mxArray* mxGetField(const mxArray* a, mwIndex index, const char* name) {
    if (!a || !a->is_struct || index != 0) return NULL;
    for (int i = 0; i < a->nfields; ++i)
        if (strcmp(a->names[i], name) == 0) return a->values[i];
    return NULL;
}

mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->m = m;
    a->n = n;
    size_t k = (m * n) != 0 ? m * n : 1;
    a->pr = (double*)calloc(k, sizeof(double));
    if (c == mxCOMPLEX) a->pi = (double*)calloc(k, sizeof(double));
    return a;
}

mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->pr[0] = v;
    return a;
}

void mxDestroyArray(mxArray* a) {
    if (!a) return;
    for (int i = 0; i < a->nfields; ++i) {
        free(a->names[i]);
        mxDestroyArray(a->values[i]);
    }
    free(a->pr);
    free(a->pi);
    free(a);
}

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err_msg, sizeof g_err_msg, fmt, ap);
    va_end(ap);
    snprintf(g_err_id, sizeof g_err_id, "%s", id ? id : "");
    if (g_jmp_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "mexErrMsgIdAndTxt outside rbshim_call: %s: %s\n", g_err_id, g_err_msg);
    abort();
}

void mexWarnMsgIdAndTxt(const char* id, const char* fmt, ...) {
    (void)fmt;
    snprintf(g_warn_id, sizeof g_warn_id, "%s", id ? id : "");
    ++g_warn_count;
}

int mexAtExit(void (*fn)(void)) {
    for (int i = 0; i < g_natexit; ++i)
        if (g_atexit[i] == fn) return 0;
    if (g_natexit < 16) g_atexit[g_natexit++] = fn;
    return 0;
}

void mexLock(void) {}

/* ---- harness helpers (ctypes) ---------------------------------------------------------------- */
typedef void (*mexfn_t)(int, mxArray**, int, const mxArray**);

/* returns 0 on success, 1 if the gateway raised mexErrMsgIdAndTxt (id/message retrievable below) */
int rbshim_call(mexfn_t fn, int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs) {
    g_err_id[0] = g_err_msg[0] = 0;
    g_jmp_armed = 1;
    if (setjmp(g_jmp)) {
        g_jmp_armed = 0;
        return 1;
    }
    fn(nlhs, plhs, nrhs, prhs);
    g_jmp_armed = 0;
    return 0;
}

const char* rbshim_last_error_id(void) { return g_err_id; }
const char* rbshim_last_error_msg(void) { return g_err_msg; }
const char* rbshim_last_warning_id(void) { return g_warn_id; }
int rbshim_warning_count(void) { return g_warn_count; }

void rbshim_run_atexit(void) {
    for (int i = g_natexit - 1; i >= 0; --i) g_atexit[i]();
    g_natexit = 0;
}

mxArray* rbshim_create_struct(void) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->is_struct = 1;
    a->m = a->n = 1;
    return a;
}

int rbshim_set_field(mxArray* s, const char* name, mxArray* value) {
    if (!s || !s->is_struct || s->nfields >= MAX_FIELDS) return 1;
    s->names[s->nfields] = strdup(name);
    s->values[s->nfields] = value;
    s->nfields++;
    return 0;
}
