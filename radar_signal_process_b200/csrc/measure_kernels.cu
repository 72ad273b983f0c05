// measure_kernels.cu -- post-CFAR measurement on the sparse detection set.
//
// Replaces CW/motionParaMeasure.m:5-87: for every flagged cell (visited in MATLAB's column-major find() order)
// the range and velocity estimates are refined by a not-a-knot cubic spline (interp1(...,'spline')) through
// 2*extraDots+1 neighbouring RDM cells, up-sampled rInterpTimes / vInterpTimes times, arg-max of the first
// maximum; the elevation comes from the monopulse sum/difference ratio and the K table.  Double precision,
// one thread per detection (O(10^2) detections: latency-trivial; the point is that the detection list never
// has to leave the device between CFAR and measurement).
#include "common.cuh"
#include "kernels.h"

namespace rb {

constexpr int kMaxSplinePts = 33;    // extraDots <= 16

// column-major compaction of the non-zero flags: count per column, scan, scatter (deterministic find() order)
__global__ void flag_count_kernel(const double* __restrict__ flags, int V, int R, int* __restrict__ col_count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int n = 0;
    for (int v = 0; v < V; ++v) n += flags[(size_t)v + (size_t)V * r] != 0.0;
    col_count[r] = n;
}

__global__ void flag_scan_kernel(int* __restrict__ col_count, int R, int* __restrict__ total) {   // single thread: R is small
    int acc = 0;
    for (int r = 0; r < R; ++r) {
        const int n = col_count[r];
        col_count[r] = acc;
        acc += n;
    }
    *total = acc;
}

// not-a-knot cubic spline through (0..n-1, y): second derivatives M by dense elimination (n <= 33)
__device__ void spline_second_derivs(const double* y, int n, double* M) {
    double A[kMaxSplinePts][kMaxSplinePts + 1];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= n; ++j) A[i][j] = 0.0;
    // not-a-knot at x_1 and x_{n-2} (unit spacing): M0 - 2 M1 + M2 = 0, M_{n-3} - 2 M_{n-2} + M_{n-1} = 0
    A[0][0] = 1.0; A[0][1] = -2.0; A[0][2] = 1.0;
    A[n - 1][n - 3] = 1.0; A[n - 1][n - 2] = -2.0; A[n - 1][n - 1] = 1.0;
    for (int i = 1; i < n - 1; ++i) {
        A[i][i - 1] = 1.0; A[i][i] = 4.0; A[i][i + 1] = 1.0;
        A[i][n] = 6.0 * (y[i + 1] - 2.0 * y[i] + y[i - 1]);
    }
    for (int c = 0; c < n; ++c) {          // Gaussian elimination with partial pivoting
        int piv = c;
        double best = fabs(A[c][c]);
        for (int i = c + 1; i < n; ++i)
            if (fabs(A[i][c]) > best) { best = fabs(A[i][c]); piv = i; }
        if (piv != c)
            for (int j = c; j <= n; ++j) { const double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        const double d = A[c][c];
        for (int i = c + 1; i < n; ++i) {
            const double f = A[i][c] / d;
            if (f != 0.0)
                for (int j = c; j <= n; ++j) A[i][j] -= f * A[c][j];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double acc = A[i][n];
        for (int j = i + 1; j < n; ++j) acc -= A[i][j] * M[j];
        M[i] = acc / A[i][i];
    }
}

// position (in cells, relative to the first node) of the first maximum of the spline sampled every 1/times
__device__ double spline_argmax(const double* y, int n, int times) {
    const int nq = (n - 1) * times + 1;
    double M[kMaxSplinePts];
    if (n >= 4) spline_second_derivs(y, n, M);
    double bestv = 0.0, bestx = 0.0;
    for (int q = 0; q < nq; ++q) {
        const double x = (double)q / (double)times;
        double val;
        if (n == 1) val = y[0];
        else if (n == 2) val = y[0] + (y[1] - y[0]) * x;
        else if (n == 3) {                   // not-a-knot with three points is the interpolating parabola
            const double a = 0.5 * (y[2] - 2.0 * y[1] + y[0]), b = y[1] - y[0] - a;
            val = y[0] + x * (b + a * x);
        } else {
            int i = (int)x;
            if (i > n - 2) i = n - 2;
            const double t = x - i, u = 1.0 - t;
            val = u * y[i] + t * y[i + 1] + ((u * u * u - u) * M[i] + (t * t * t - t) * M[i + 1]) / 6.0;
        }
        if (q == 0 || val > bestv) { bestv = val; bestx = x; }
    }
    return bestx;
}

struct MeasureParams {
    const double *sum, *diff, *flags, *rScale, *vScale, *kValues;
    int V, R, extra, rTimes, vTimes, n0, kRows;
    double deltaR, deltaV, beamPosNum, beamAngleStep, eleComp, eleSysErr;
    int freInd;
    const int* col_start;
    double *out_r, *out_v, *out_e;
    int* err_flag;
};

__global__ void measure_kernel(const MeasureParams p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per range column keeps find() order
    if (r >= p.R) return;
    int slot = p.col_start[r];
    const int k = p.extra, n = 2 * k + 1;
    for (int v = 0; v < p.V; ++v) {
        if (p.flags[(size_t)v + (size_t)p.V * r] == 0.0) continue;
        double y[kMaxSplinePts];
        // ---- range (motionParaMeasure.m:22-43), 1-based cell arithmetic
        int c0 = (r + 1) - k;
        if (c0 <= 0) c0 = 1;
        if (c0 + 2 * k > p.R) c0 = p.R - 2 * k;
        if (c0 < 1) { *p.err_flag = 1; return; }
        for (int i = 0; i < n; ++i) y[i] = p.sum[(size_t)v + (size_t)p.V * (c0 - 1 + i)];
        const double rCellMax = c0 + spline_argmax(y, n, p.rTimes);
        const double rEst = p.rScale[r] + (rCellMax - (r + 1)) * p.deltaR;
        // ---- velocity (:49-70)
        int v0 = (v + 1) - k;
        if (v0 <= p.n0 + 1) v0 = p.n0 + 2;
        if (v0 + 2 * k > p.V - p.n0) v0 = p.V - p.n0 - 2 * k;
        if (v0 < 1 || v0 + 2 * k > p.V) { *p.err_flag = 1; return; }
        for (int i = 0; i < n; ++i) y[i] = p.sum[(size_t)(v0 - 1 + i) + (size_t)p.V * r];
        const double vCellMax = v0 + spline_argmax(y, n, p.vTimes);
        const int fx = (int)vCellMax;
        const double vEst = p.vScale[fx - 1] - (vCellMax - fx) * p.deltaV;
        // ---- elevation (:76-79)
        const double ratio = p.diff[(size_t)v + (size_t)p.V * r] / p.sum[(size_t)v + (size_t)p.V * r];
        const double kv = p.kValues[(size_t)p.freInd + (size_t)p.kRows * (int)p.beamPosNum];
        const double eEst = p.beamPosNum * p.beamAngleStep + 2.5 - ratio * kv + p.eleComp + p.eleSysErr;
        p.out_r[slot] = rEst;
        p.out_v[slot] = vEst;
        p.out_e[slot] = eEst;
        ++slot;
    }
}

cudaError_t launch_flag_compaction(const double* flags, int V, int R, int* col_start, int* total, cudaStream_t st) {
    flag_count_kernel<<<(R + 127) / 128, 128, 0, st>>>(flags, V, R, col_start);
    flag_scan_kernel<<<1, 1, 0, st>>>(col_start, R, total);
    return cudaGetLastError();
}

cudaError_t launch_measure(const double* sum, const double* diff, const double* flags, const double* rScale, const double* vScale,
                           const double* kValues, int kRows, int V, int R, int extra, int rTimes, int vTimes, int n0, double deltaR,
                           double deltaV, double beamPosNum, double beamAngleStep, int freInd, double eleComp, double eleSysErr,
                           const int* col_start, double* out_r, double* out_v, double* out_e, int* err_flag, cudaStream_t st) {
    MeasureParams p;
    p.sum = sum; p.diff = diff; p.flags = flags; p.rScale = rScale; p.vScale = vScale; p.kValues = kValues;
    p.V = V; p.R = R; p.extra = extra; p.rTimes = rTimes; p.vTimes = vTimes; p.n0 = n0; p.kRows = kRows;
    p.deltaR = deltaR; p.deltaV = deltaV; p.beamPosNum = beamPosNum; p.beamAngleStep = beamAngleStep;
    p.eleComp = eleComp; p.eleSysErr = eleSysErr; p.freInd = freInd;
    p.col_start = col_start; p.out_r = out_r; p.out_v = out_v; p.out_e = out_e; p.err_flag = err_flag;
    measure_kernel<<<(R + 63) / 64, 64, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace rb
