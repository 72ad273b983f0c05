/* mex.h -- minimal stand-in for the MATLAB / GNU Octave MEX API (Octave-compatible subset only).
 *
 * Neither MATLAB nor Octave exists in the build image, so the gateways in mex/ are compiled against
 * this header and executed through mex/shim/mex_shim.c by the tests.  Against a real interpreter the
 * same gateway sources compile with `mex` / `mkoctfile --mex` and the interpreter's own mex.h; only
 * the functions declared here are used (SURVEY.md section 8b "gateway rules").
 */
#ifndef RB200_MEX_SHIM_H
#define RB200_MEX_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

/* queries */
double* mxGetPr(const mxArray* a);
double* mxGetPi(const mxArray* a);
size_t mxGetM(const mxArray* a);
size_t mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsStruct(const mxArray* a);
int mxIsEmpty(const mxArray* a);
double mxGetScalar(const mxArray* a);
mxArray* mxGetField(const mxArray* a, mwIndex index, const char* name);

/* creation (owned by the interpreter once returned through plhs) */
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
void mxDestroyArray(mxArray* a);

/* interpreter services */
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);   /* does not return (longjmp) */
void mexWarnMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexAtExit(void (*fn)(void));
void mexLock(void);

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

#ifdef __cplusplus
}
#endif
#endif
