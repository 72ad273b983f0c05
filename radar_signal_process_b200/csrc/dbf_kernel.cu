// dbf_kernel.cu -- digital beam forming fused with the int16 DDC unpack.
//
// Replaces FrameDataRead_xzr.m:150-158: sig_C = I + jQ per channel, then
//     current_sig_data_DBF = sig_C * DBF_coeffs_data_C.'      (n x 16) . (16 x 13), non-conjugate transpose
// i.e. beam b of range cell r is sum_ch sig[r, ch] * W[b, ch].
//
// One thread owns one range cell of one PRT: it reads the cell's interleaved channels with 128-bit loads,
// keeps them in registers, and produces every beam with W broadcast from shared memory.  Output is the planar
// layout the pulse-compression kernel consumes: float2 [cpi][beam][prt][range] (range-contiguous stores).
#include "common.cuh"
#include "kernels.h"

namespace rb {

template <int CH>
__global__ void __launch_bounds__(128)
dbf_kernel(const int* __restrict__ raw, float2* __restrict__ out, const float2* __restrict__ W, int n_beams, int n_ch, int P, int R) {
    extern __shared__ float2 w_sm[];              // [beam][ch]
    for (int i = threadIdx.x; i < n_beams * n_ch; i += blockDim.x) w_sm[i] = __ldg(W + i);
    __syncthreads();
    const int g = blockIdx.y;
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= R) return;
    const int cpi = g / P, prt = g - cpi * P;
    const size_t cell = (size_t)g * R + r;
    if (CH > 0) {
        float2 x[CH > 0 ? CH : 1];
        const int4* src = reinterpret_cast<const int4*>(raw + cell * CH);
#pragma unroll
        for (int q = 0; q < CH / 4; ++q) {
            const int4 w = __ldg(src + q);
            const int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) x[4 * q + e] = make_float2((float)(short)(ww[e] & 0xffff), (float)(ww[e] >> 16));
        }
        for (int b = 0; b < n_beams; ++b) {
            const float2* wb = w_sm + b * CH;
            float ax = 0.f, ay = 0.f;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const float2 w = wb[ch];
                ax = fmaf(x[ch].x, w.x, ax); ax = fmaf(-x[ch].y, w.y, ax);
                ay = fmaf(x[ch].x, w.y, ay); ay = fmaf(x[ch].y, w.x, ay);
            }
            out[(((size_t)cpi * n_beams + b) * P + prt) * R + r] = make_float2(ax, ay);
        }
    } else {
        for (int b = 0; b < n_beams; ++b) {
            const float2* wb = w_sm + b * n_ch;
            float ax = 0.f, ay = 0.f;
            for (int ch = 0; ch < n_ch; ++ch) {
                const int wv = __ldg(raw + cell * n_ch + ch);
                const float xr = (float)(short)(wv & 0xffff), xi = (float)(wv >> 16);
                const float2 w = wb[ch];
                ax = fmaf(xr, w.x, ax); ax = fmaf(-xi, w.y, ax);
                ay = fmaf(xr, w.y, ay); ay = fmaf(xi, w.x, ay);
            }
            out[(((size_t)cpi * n_beams + b) * P + prt) * R + r] = make_float2(ax, ay);
        }
    }
}

// DBF-type (data_type 2) payload: rows of row_bytes bytes per range sample, 3-byte little-endian words,
// value > 2^23 -> value - 2^24 (FrameDataRead_xzr.m:130-135,163).  One thread per output complex value.
// out: float2 planar [beam][prt][range].
__global__ void unpack_dbf24_kernel(const uint8_t* __restrict__ bytes, float2* __restrict__ out, int n_prt, int n, int ncol,
                                    int row_bytes, size_t prt_bytes) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n_prt * n * ncol;
    if (i >= total) return;
    const int r = (int)(i % n);
    const size_t t1 = i / n;
    const int prt = (int)(t1 % n_prt);
    const int b = (int)(t1 / n_prt);
    const uint8_t* q = bytes + (size_t)prt * prt_bytes + (size_t)r * row_bytes + (size_t)b * 6;
    int vi = (int)q[0] | ((int)q[1] << 8) | ((int)q[2] << 16);
    int vq = (int)q[3] | ((int)q[4] << 8) | ((int)q[5] << 16);
    if (vi > (1 << 23)) vi -= (1 << 24);
    if (vq > (1 << 23)) vq -= (1 << 24);
    out[i] = make_float2((float)vi, (float)vq);       // |v| <= 2^23: exact in fp32
}

cudaError_t launch_unpack_dbf24(const uint8_t* bytes, float2* out, int n_prt, int n, int ncol, int row_bytes, size_t prt_bytes, cudaStream_t st) {
    const size_t total = (size_t)n_prt * n * ncol;
    if (total == 0) return cudaSuccess;
    unpack_dbf24_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bytes, out, n_prt, n, ncol, row_bytes, prt_bytes);
    return cudaGetLastError();
}

cudaError_t launch_dbf(const int16_t* raw, float2* out, const float2* W, int n_beams, int n_ch, int n_groups, int P, int R, cudaStream_t st) {
    if (n_groups <= 0 || n_beams <= 0) return cudaSuccess;
    if (n_groups > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid((R + 127) / 128, n_groups, 1);
    const size_t smem = (size_t)n_beams * n_ch * sizeof(float2);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    const int* rw = reinterpret_cast<const int*>(raw);
    if (n_ch == 16 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0) dbf_kernel<16><<<grid, 128, smem, st>>>(rw, out, W, n_beams, n_ch, P, R);
    else dbf_kernel<0><<<grid, 128, smem, st>>>(rw, out, W, n_beams, n_ch, P, R);
    return cudaGetLastError();
}

}  // namespace rb
