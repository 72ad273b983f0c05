// mtd_kernels.cu -- K2: [MTI] + Kaiser window + slow-time FFT + fftshift + |.| + zero-velocity mask.
//
// Replaces the per-range-cell loop of MP/fun_Process_MTD.m:20-30 (window .* column, fft, fftshift,
// abs) fused with MP/fun_0v_pressing.m:4-6 (rows zeroed) and, optionally, the 30-pulse canceller of
// MP/fun_Process_MTI.m:20-22.  fftshift is pure output indexing: Doppler bin m lands on row
// (m + floor(P/2)) mod P.
//
// Input  : pulse-compressed samples, float2 planar [slab][prt][range]  (slab = cpi*lanes + lane)
// Output : RDM magnitudes, float [slab][v][range]
//
// Fast path (P = R^2: 64 = 8*8, 256 = 16*16): one CTA owns TR adjacent range cells of one slab; a
// thread holds R slow-time samples of one range cell (stride R apart), runs a radix-R butterfly in
// registers, exchanges through shared memory once (the corner turn), runs the second radix-R
// butterfly and stores |X|.  All global accesses are range-contiguous (256 B per warp load, 128 B per
// warp store).
// Generic path (any P): out-of-place Stockham with one thread per output element and stage radices
// chosen on the host (any factorisation, prime factors included), twiddles from a double-computed
// table -- used by the parity configurations (P = 8, 1536, ...) and the MATLAB-layout entry points.
#include "common.cuh"
#include "radix.cuh"
#include "kernels.h"

namespace rb {

__device__ __forceinline__ float2 mtd_load(const float2* __restrict__ col, int prt, int P, int ld, int lag) {
    // col points at (prt 0, this range cell); MTI: x[p+lag] - x[p], last `lag` rows are zero
    if (lag == 0) return __ldg(col + (size_t)prt * ld);
    if (prt >= P - lag) return make_float2(0.f, 0.f);
    const float2 a = __ldg(col + (size_t)(prt + lag) * ld);
    const float2 b = __ldg(col + (size_t)prt * ld);
    return make_float2(a.x - b.x, a.y - b.y);
}

template <int R, int TR>
__global__ void __launch_bounds__(TR * R)
mtd_fast_kernel(const MtdParams p) {
    constexpr int P = R * R;
    extern __shared__ float2 sm[];   // [P][TR]
    const int rl = threadIdx.x % TR;
    const int u = threadIdx.x / TR;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * TR + rl;
    const bool ok = r < p.cols;
    const float2* col = p.in + (size_t)slab * P * p.in_ld + r;

    float2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int prt = u + j * R;
        float2 x = make_float2(0.f, 0.f);
        if (ok) x = mtd_load(col, prt, P, p.in_ld, p.mti_lag);
        v[j] = cscale(x, __ldg(p.window + prt));
    }
    Dft<R, -1>::run(v);
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cmul(v[k], __ldg(p.tw + u * k));
#pragma unroll
    for (int k = 0; k < R; ++k) sm[(u + k * R) * TR + rl] = v[k];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = sm[(u * R + j) * TR + rl];
    Dft<R, -1>::run(v);
    if (ok) {
        float* out = p.out + (size_t)slab * P * p.out_ld + r;
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
            const int m = u + R * k1;                  // Doppler bin
            const int row = (m + P / 2) & (P - 1);     // fftshift
            float mag = sqrtf(v[k1].x * v[k1].x + v[k1].y * v[k1].y);
            if (row >= p.zv_lo && row <= p.zv_hi) mag = 0.f;
            out[(size_t)row * p.out_ld] = mag;
        }
    }
}

// Generic Stockham: blockDim = (TR, NY); shared = 2 * P * TR float2.
__global__ void mtd_generic_kernel(const MtdParams p) {
    extern __shared__ float2 sm[];
    const int TR = blockDim.x;
    const int P = p.P;
    float2* a = sm;
    float2* b = sm + (size_t)P * TR;
    const int rl = threadIdx.x;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * TR + rl;
    const bool ok = r < p.cols;
    const float2* col = p.in + (size_t)slab * P * p.in_ld + r;
    for (int prt = threadIdx.y; prt < P; prt += blockDim.y) {
        float2 x = make_float2(0.f, 0.f);
        if (ok) x = mtd_load(col, prt, P, p.in_ld, p.mti_lag);
        a[prt * TR + rl] = cscale(x, __ldg(p.window + prt));
    }
    __syncthreads();
    int Ns = 1;
    for (int s = 0; s < p.n_stages; ++s) {
        const int Rr = p.radix[s];
        const int span = Ns * Rr;
        const int leg = P / Rr;
        for (int o = threadIdx.y; o < P; o += blockDim.y) {
            const int m = o % span;
            const int j = (o / span) * Ns + (m % Ns);
            const int step = (int)(((long long)m * (P / span)) % P);
            int idx = 0;
            float ax = 0.f, ay = 0.f;
            for (int t = 0; t < Rr; ++t) {
                const float2 x = a[(j + t * leg) * TR + rl];
                const float2 w = __ldg(p.tw + idx);
                ax = fmaf(x.x, w.x, ax); ax = fmaf(-x.y, w.y, ax);
                ay = fmaf(x.x, w.y, ay); ay = fmaf(x.y, w.x, ay);
                idx += step;
                if (idx >= P) idx -= P;
            }
            b[o * TR + rl] = make_float2(ax, ay);
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
        Ns = span;
    }
    if (ok) {
        float* out = p.out + (size_t)slab * P * p.out_ld + r;
        const int half = P / 2;
        for (int m = threadIdx.y; m < P; m += blockDim.y) {
            int row = m + half;
            if (row >= P) row -= P;
            const float2 x = a[m * TR + rl];
            float mag = hypotf(x.x, x.y);
            if (row >= p.zv_lo && row <= p.zv_hi) mag = 0.f;
            out[(size_t)row * p.out_ld] = mag;
        }
    }
}

bool mtd_has_fast_path(int P) { return P == 64 || P == 256; }
int mtd_generic_max_p() { return 12288; }

template <int R, int TR>
static cudaError_t launch_fast(const MtdParams& p, int n_slabs, cudaStream_t st) {
    constexpr int P = R * R;
    const size_t smem = (size_t)P * TR * sizeof(float2);
    static size_t configured[64] = {};
    cudaError_t ce = ensure_dynamic_smem(mtd_fast_kernel<R, TR>, smem, configured);
    if (ce != cudaSuccess) return ce;
    dim3 grid((p.cols + TR - 1) / TR, n_slabs, 1);
    if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
    mtd_fast_kernel<R, TR><<<grid, TR * R, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_mtd(const MtdParams& p, int n_slabs, cudaStream_t st) {
    if (p.cols <= 0 || n_slabs <= 0) return cudaSuccess;
    if (p.P == 64) return launch_fast<8, 64>(p, n_slabs, st);
    if (p.P == 256) return launch_fast<16, 32>(p, n_slabs, st);
    // generic
    int TR = 32;
    while (TR > 1 && (size_t)2 * p.P * TR * sizeof(float2) > 96 * 1024) TR >>= 1;
    const size_t smem = (size_t)2 * p.P * TR * sizeof(float2);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static size_t configured[64] = {};
    cudaError_t ce = ensure_dynamic_smem(mtd_generic_kernel, smem, configured);
    if (ce != cudaSuccess) return ce;
    dim3 block(TR, 256 / TR, 1);
    dim3 grid((p.cols + TR - 1) / TR, n_slabs, 1);
    if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
    mtd_generic_kernel<<<grid, block, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace rb
