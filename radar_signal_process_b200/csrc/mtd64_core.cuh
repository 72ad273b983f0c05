// mtd64_core.cuh -- the register-resident Doppler column (64-point FFT, |.|, zero-velocity mask, velocity
// CA-CFAR, hit compaction) shared by mtd64_kernel.cu and chain64_kernel.cu.
#pragma once
#include "common.cuh"
#include "radix.cuh"
#include "tw64.cuh"
#include "../../include/radar_b200.h"

namespace rb {

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#ifndef RB200_MTD64_MINB
#define RB200_MTD64_MINB 3
#endif

// The column work shared by both kernel variants: v[] holds the 64 windowed slow-time samples of range cell r.
template <int REF, int GUARD, int N0, bool CFAR>
__device__ __forceinline__ void mtd64_column(float2 (&v)[64], const Mtd64Params& p, int slab, int r, bool ok) {
    constexpr int P = 64;
    // ---- 64-point DIF: step 1, radix-8 over j for every q (elements q + 8j), twiddle w64^(q*k0) ----
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float2 a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = v[q + 8 * j];
        dft8<-1>(a);
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) {
            const int m = (q * k0) & 63;
            v[q + 8 * k0] = (m == 0) ? a[k0] : cmul(a[k0], make_float2(kCos64[m], -kSin64[m]));
        }
    }
    // ---- step 2: radix-8 over q for every k0 (elements 8*k0 .. 8*k0+7) -> X[k0 + 8*k1] at 8*k0 + k1 ----
    float mag[P];   // indexed by output row (fftshifted)
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) {
        float2 a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = v[q + 8 * k0];
        dft8<-1>(a);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const int row = (k0 + 8 * k1 + P / 2) & (P - 1);
            mag[row] = fast_sqrt(a[k1].x * a[k1].x + a[k1].y * a[k1].y) * p.keep[row];
        }
    }
    if (ok) {
        float* out = p.out + (size_t)slab * P * p.out_ld + r;
#pragma unroll
        for (int row = 0; row < P; ++row) out[(size_t)row * p.out_ld] = mag[row];
    }
    if (!CFAR) return;

    // ---- velocity-axis CA-CFAR on the register column (tested rows N0+1 .. 63-N0) ----
    constexpr int NV = P - 2 * N0 - 1;
    static_assert(!CFAR || NV >= 2 * (REF + GUARD), "velocity axis shorter than 2*(ref+guard)");
    // window sums are taken directly from the REF cells (compile-time indices, registers only): a running prefix over the
    // whole column would quantise the sums behind a strong target to the ulp of the peak
    unsigned long long hits = 0ull;
#pragma unroll
    for (int y = 0; y < NV; ++y) {
        const int l1 = y - GUARD - REF;
        const int r1 = y + GUARD + 1;
        const bool okL = l1 >= 0;
        const bool okR = r1 + REF - 1 <= NV - 1;
        float sl = 0.f, sr = 0.f;
#pragma unroll
        for (int j = 0; j < REF; ++j) {
            if (okL) sl += mag[N0 + 1 + (okL ? l1 + j : 0)];
            if (okR) sr += mag[N0 + 1 + (okR ? r1 + j : 0)];
        }
        const float a = okL ? sl : sr;
        const float b = okR ? sr : sl;
        const float mu = p.meth_v == 0 ? fmaxf(a, b) : fminf(a, b);
        if (mag[N0 + 1 + y] >= mu * p.tv_over_ref) hits |= 1ull << (N0 + 1 + y);
    }
    {
        int slo, shi;
        if (!ok || !cfar_seg_of(p.segs, r, p.cols, &slo, &shi)) hits = 0ull;
    }
    if (ok) p.colmask[(size_t)slab * p.cols_ld + r] = hits;
    // ---- compaction: one atomic per warp ----
    if (!__any_sync(0xffffffffu, hits != 0ull)) return;
    const int lane = threadIdx.x & 31;
    const int n = __popcll(hits);
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31) base = atomicAdd(p.det_count, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    int slot = base + incl - n;
    unsigned long long h = hits;
    while (h) {
        const int row = __ffsll((long long)h) - 1;
        h &= h - 1;
        if (slot < p.max_det) {
            rb200_det d;
            d.cpi = (uint32_t)(p.cpi0 + slab / p.n_lanes);
            d.r = (uint32_t)r;
            d.v = (uint16_t)row;
            d.lane = (uint8_t)(slab % p.n_lanes);
            d.kind = RB200_DET_V;
            // mag[] is register-resident with static indexing only: re-read the stored magnitude
            d.amp = p.out[((size_t)slab * P + row) * p.out_ld + r];
            reinterpret_cast<rb200_det*>(p.dets)[slot] = d;
        }
        ++slot;
    }
}

}  // namespace rb
