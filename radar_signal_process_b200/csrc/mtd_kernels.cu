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
#include "cfar_core.cuh"
#include "pc_core.cuh"
#include "tmap.h"
#include <algorithm>
#include "../../include/radar_b200.h"

namespace rb {

__device__ __forceinline__ float2 mtd_load(const float2* __restrict__ col, int prt, int P, int ld, int lag) {
    // col points at (prt 0, this range cell); MTI: x[p+lag] - x[p], last `lag` rows are zero
    if (lag == 0) return __ldg(col + (size_t)prt * ld);
    if (prt >= P - lag) return make_float2(0.f, 0.f);
    const float2 a = __ldg(col + (size_t)(prt + lag) * ld);
    const float2 b = __ldg(col + (size_t)prt * ld);
    return make_float2(a.x - b.x, a.y - b.y);
}

__device__ __forceinline__ float mtd_fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // <= 2 ulp; the parity bar is 1e-4 relative
    return r;
}

// Everything after the loads, shared by the one-tile-per-CTA kernel and the persistent TMA-fed one: window, radix-R butterflies,
// corner turn through `sm` ([P][TR] float2), second butterflies, |.|, zero-velocity rows, RDM stores and (CF != 0) the fused
// velocity-axis CFAR on the magnitude tile that re-uses `sm`.  Contains CTA-wide barriers: every thread of the CTA calls it.
// barrier of the thread group that works on one tile: the whole CTA, or one 512-thread half of it (named barrier)
struct MtdCtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct MtdHalfSync {
    int id;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory"); }
};
struct MtdNoHook {
    __device__ __forceinline__ void operator()() const {}
};
template <int R, int TR, int CF, bool PRE_SYNC, class Sync, class Hook>
__device__ __forceinline__ void mtd_fast_body(float2 (&v)[R], const MtdParams& p, float2* sm, const float* win_sm, const float2* tw_sm,
                                              const int slab, const int r, const bool ok, const int u, const int rl, const Sync sync,
                                              const Hook after_first_sync) {
    constexpr int P = R * R;
    {
        const float* wu = win_sm + u;
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = cscale(v[j], wu[j * R]);
    }
    Dft<R, -1>::run(v);
    {
        const float2* twu = tw_sm + u * R;
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmul(v[k], twu[k]);
    }
    // persistent callers: `sm` still holds the magnitude tile of the group's previous item until everybody has left its CFAR;
    // waiting HERE rather than at the end of that item lets the early warps run this item's loads and first butterflies
    if (PRE_SYNC) sync();
#pragma unroll
    for (int k = 0; k < R; ++k) sm[(u + k * R) * TR + rl] = v[k];
    sync();
    after_first_sync();
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = sm[(u * R + j) * TR + rl];
    Dft<R, -1>::run(v);
    // Output rows of this thread after fftshift: Doppler bin u + R*k1 lands on row u + P/2 + R*k1 for k1 < R/2 and on
    // row u + R*(k1 - R/2) for the rest, i.e. two runs of constant stride -> stepped byte addresses, immediate offsets.
    // zrow0: first row >= zv_lo that is congruent to u (mod R); rows of this thread are zeroed iff zrow0 <= zv_hi
    // (u is warp-uniform, so the common "nothing to zero" case skips every per-row test).
    const int zrow0 = p.zv_lo + ((u - p.zv_lo) % R + R) % R;
    const bool zany = p.zv_lo <= p.zv_hi && zrow0 <= p.zv_hi;
    const unsigned long long rowb = (unsigned long long)p.out_ld * sizeof(float);
    const unsigned long long o_lo = reinterpret_cast<unsigned long long>(p.out + (size_t)slab * P * p.out_ld + r) + u * rowb;   // row u
    const unsigned long long o_hi = o_lo + (P / 2) * rowb;                                                                     // row u + P/2
    const unsigned long long ostep = R * rowb;
    // The magnitude tile [P][TR] re-uses the exchange buffer (2*P*TR floats) once everybody has read it; it sits
    // P/2 rows into the buffer so that window reads up to P/2 rows outside the tile stay inside the allocation
    // (those values are never used: the edge rule replaces the side that does not fit).
    float* mag_sm = reinterpret_cast<float*>(sm) + (P / 2) * TR;
    if (CF != 0) sync();
    float* mag_u = mag_sm + u * TR + rl;               // row u of this thread's column
    if (CF == 0 && !ok) return;
    if (zany) {                                            // warp-uniform (u is): only the few warps that own a zeroed row test
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
            const int roff = k1 < R / 2 ? P / 2 + R * k1 : R * (k1 - R / 2);     // row - u, compile-time
            float mag = mtd_fast_sqrt(v[k1].x * v[k1].x + v[k1].y * v[k1].y);
            if (u + roff >= p.zv_lo && u + roff <= p.zv_hi) mag = 0.f;
            const unsigned long long oa = (k1 < R / 2 ? o_hi + k1 * ostep : o_lo + (k1 - R / 2) * ostep);
            if (ok) *reinterpret_cast<float*>(oa) = mag;
            if (CF != 0) mag_u[roff * TR] = mag;
        }
    } else {
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
            const int roff = k1 < R / 2 ? P / 2 + R * k1 : R * (k1 - R / 2);
            const float mag = mtd_fast_sqrt(v[k1].x * v[k1].x + v[k1].y * v[k1].y);
            const unsigned long long oa = (k1 < R / 2 ? o_hi + k1 * ostep : o_lo + (k1 - R / 2) * ostep);
            if (ok) *reinterpret_cast<float*>(oa) = mag;
            if (CF != 0) mag_u[roff * TR] = mag;
        }
    }
    if (CF == 0) return;
    // ---- fused velocity-axis CA-CFAR (CW/executeCFAR.m:28, CW/Function_CFAR1D_sub.m:17-69) on the tile ----
    // Thread (u, rl) decides rows [u*R, (u+1)*R) of column rl; the reference windows come from shared memory.
    sync();
    const int nv = p.cf.v_hi - p.cf.v_lo;
    const int ref = CF == 2 ? 5 : p.cf.ref_v;
    const int guard = CF == 2 ? 7 : p.cf.guard_v;
    const int H = ref + guard;
    const int y0 = u * R - p.cf.v_lo;                  // axis position of this thread's first row (may be negative)
    const float* mcol = mag_sm + p.cf.v_lo * TR + rl;  // mcol[y*TR] = x[v_lo + y][r]
    float ml[R], mr[R], xc[R];
    if (CF == 2) {
        constexpr int HH = 12, W = R + 2 * HH;
        float w[W];
#pragma unroll
        for (int k = 0; k < W; ++k) w[k] = mcol[(y0 - HH + k) * TR];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            ml[i] = ((((w[i] + w[i + 1]) + w[i + 2]) + w[i + 3]) + w[i + 4]);
            mr[i] = ((((w[i + 20] + w[i + 21]) + w[i + 22]) + w[i + 23]) + w[i + 24]);
            xc[i] = w[i + HH];
        }
    } else {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            ml[i] = 0.f;
            mr[i] = 0.f;
            xc[i] = mcol[(y0 + i) * TR];
        }
        const float* pl = mcol + (y0 - H) * TR;
        const float* pr = mcol + (y0 + guard + 1) * TR;
#pragma unroll 1
        for (int j = 0; j < ref; ++j) {                // R independent accumulations per side in flight
#pragma unroll
            for (int i = 0; i < R; ++i) {
                ml[i] += pl[i * TR];
                mr[i] += pr[i * TR];
            }
            pl += TR;
            pr += TR;
        }
    }
    const float inv_ref = 1.f / (float)ref;
    // decisions of this thread's R rows as a bit mask; `tested` (warp-uniform) marks the rows inside [v_lo, v_hi)
    unsigned hits = 0u, tested = 0u;
    const bool interior = y0 - H >= 0 && y0 + R - 1 + H <= nv - 1;      // both windows fit for all R rows (most threads)
    if (interior) {
        tested = (1u << R) - 1u;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const float a = ml[i] * inv_ref, b = mr[i] * inv_ref;
            const float mu = p.cf.meth_v == 0 ? fmaxf(a, b) : fminf(a, b);
            hits |= (xc[i] >= mu * p.t_v ? 1u : 0u) << i;
        }
    } else {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int y = y0 + i;
            if (y < 0 || y >= nv) continue;
            const bool okL = y - H >= 0;
            const bool okR = y + H <= nv - 1;
            if (!okL && !okR) {                        // MATLAB: index exceeds array bounds
                if (threadIdx.x == 0) *p.err_flag = 1;
                continue;
            }
            tested |= 1u << i;
            const float a = (okL ? ml[i] : mr[i]) * inv_ref;
            const float b = (okR ? mr[i] : ml[i]) * inv_ref;
            const float mu = p.cf.meth_v == 0 ? fmaxf(a, b) : fminf(a, b);
            hits |= (xc[i] >= mu * p.t_v ? 1u : 0u) << i;
        }
    }
    {
        int slo, shi;
        if (!ok || !cfar_seg_of(p.cf.segs, r, p.cf.R, &slo, &shi)) hits = 0u;      // columns outside every range segment
    }
    const int Rw = (p.cf.R + 31) / 32;
    const int lane = threadIdx.x & 31;
    uint32_t* vm = p.vmask + ((size_t)slab * p.cf.V + u * R) * Rw + (r >> 5);   // a warp covers 32 consecutive columns
    const bool vm_ok = lane == 0 && (r >> 5) < Rw;
    const bool any = __any_sync(0xffffffffu, hits != 0u);
    if (!any) {                                        // the common case: no detection in these 32 columns x R rows
        if (vm_ok) {
#pragma unroll
            for (int i = 0; i < R; ++i)
                if ((tested >> i) & 1u) vm[i * Rw] = 0u;
        }
        return;
    }
    const uint32_t det_hdr = (uint32_t)(slab % p.cf.n_lanes) << 16 |
                             (uint32_t)(p.cf.range_stage ? RB200_DET_V : (RB200_DET_V | RB200_DET_2D)) << 24;
    const uint32_t det_cpi = (uint32_t)(p.cf.cpi0 + slab / p.cf.n_lanes);
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (!((tested >> i) & 1u)) continue;           // warp-uniform: u is warp-uniform for TR >= 32
        const bool hit = (hits >> i) & 1u;
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        if (vm_ok) vm[i * Rw] = ball;
        if (ball == 0) continue;
        int basei = 0;
        if (lane == 0) basei = atomicAdd(p.det_count, __popc(ball));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (hit) {
            const int slot = basei + __popc(ball & ((1u << lane) - 1u));
            if (slot < p.cf.max_det) {
                // rb200_det {u32 cpi, u32 r, u16 v, u8 lane, u8 kind, f32 amp} written as one 16-byte store
                uint4 d;
                d.x = det_cpi;
                d.y = (uint32_t)r;
                d.z = (uint32_t)(u * R + i) | det_hdr;
                d.w = __float_as_uint(xc[i]);
                reinterpret_cast<uint4*>(p.dets)[slot] = d;
            }
        }
    }
}


// CF: 0 = no CFAR, 1 = fused velocity CFAR with run-time (ref, guard), 2 = fused with the reference's default
// protection/reference cells (5 reference, 7 guard: CW/main_cfar.m:147-154) as compile-time constants.
template <int R, int TR, bool MTI, int CF>
__global__ void __launch_bounds__(TR * R, 2)
mtd_fast_kernel(const MtdParams p) {
    constexpr int P = R * R;
    extern __shared__ float2 sm[];   // [P][TR]
    const int rl = threadIdx.x % TR;
    const int u = threadIdx.x / TR;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * TR + rl;
    const bool ok = r < p.cols;
    // one 64-bit byte address per thread, then a constant byte stride: pulse u + j*R sits j*stepb bytes further on
    // (kept as integers so that the compiler steps the address instead of re-deriving it from an element index)
    unsigned long long a0 = reinterpret_cast<unsigned long long>(p.in + (size_t)slab * P * p.in_ld + (ok ? r : 0) + (size_t)u * p.in_ld);
    const unsigned long long stepb = (unsigned long long)R * p.in_ld * sizeof(float2);

    // window and first-stage twiddles of the CTA in shared memory (every later access is base + immediate): win_sm[prt],
    // tw_sm[u][k] = w_P^(u k); filled while the data loads below are in flight
    __shared__ float win_sm[P];
    __shared__ float2 tw_sm[P];
    float2 v[R];
    if (MTI) {
        // x[p + lag] - x[p], the last `lag` pulses are zero (MP/fun_Process_MTI.m:20-22)
        unsigned long long a1 = a0 + (unsigned long long)p.mti_lag * p.in_ld * sizeof(float2);
        const int last = P - p.mti_lag - u;            // pulse u + j*R is kept iff j*R < last
#pragma unroll
        for (int j = 0; j < R; ++j) {
            float2 x = make_float2(0.f, 0.f);
            if (j * R < last) x = csub(__ldg(reinterpret_cast<const float2*>(a1)), __ldg(reinterpret_cast<const float2*>(a0)));
            v[j] = x;
            a0 += stepb;
            a1 += stepb;
        }
    } else {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            v[j] = __ldg(reinterpret_cast<const float2*>(a0));
            a0 += stepb;
        }
    }
    for (int i = threadIdx.x; i < P; i += TR * R) {
        win_sm[i] = __ldg(p.window + i);
        tw_sm[i] = __ldg(p.tw + (i / R) * (i % R));
    }
    __syncthreads();
    mtd_fast_body<R, TR, CF, false>(v, p, sm, win_sm, tw_sm, slab, r, ok, u, rl, MtdCtaSync(), MtdNoHook());
}

// Persistent TMA-fed variant for P = 256 (the DBF-mode CPI): one CTA of 1 024 threads per SM, two HALVES of 512 threads that
// each walk over their own (slab, 32-column) tiles with the arithmetic of mtd_fast_kernel<16, 32, ...> (mtd_fast_body, barriers
// = named barriers of the half).  The 256 x 32 input tile of an item (64 KB) arrives through ONE tensor-map TMA load (3-D map
// over [slab][prt][2 x range] floats; columns past `cols` are zero-filled by the copy engine) into a staging buffer the two
// halves use in turn: as soon as a half has pulled its samples into registers (its first barrier) it starts the load of the
// OTHER half's next tile, so that every tile is in flight while both halves run butterflies and CFAR -- 32 warps stay in
// the arithmetic phases instead of alternating with exposed load phases.  The MTI difference x[p + lag] - x[p] reads both pulses
// from the staged tile (the one-tile-per-CTA kernel fetches every row twice).  Generic-proxy reads of the staging buffer are
// ordered before the async-proxy refill by fence.proxy.async + the half's barrier.
namespace m256 {
constexpr int kR = 16, kTR = 32, kP = 256;
constexpr int kHalf = kR * kTR;                      // 512 threads work on one tile
constexpr int kThreads = 2 * kHalf;
constexpr int kTileBytes = kP * kTR * 8;             // 65 536
constexpr int kSmemBytes = 3 * kTileBytes + 128;     // the staging tile + one exchange / magnitude buffer per half
}  // namespace m256

__device__ __forceinline__ void mtd_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// starts the load of `item` into the staging tile; called by one thread once the previous user has released the tile
struct Mtd256Issue {
    const CUtensorMap* map;
    float2* stage;
    uint64_t* bar;
    int item, tiles_per_slab;
    bool on;
    __device__ __forceinline__ void operator()() const {
        if (!on) return;
        const int slab = item / tiles_per_slab;
        const int c0 = (item - slab * tiles_per_slab) * m256::kTR;
        mbar_expect_tx(bar, (uint32_t)m256::kTileBytes);
        mtd_tma_load_3d(stage, map, 2 * c0, 0, slab, bar);
    }
};

template <bool MTI, int CF>
__global__ void __launch_bounds__(m256::kThreads, 1)
mtd256_tma_kernel(const __grid_constant__ MtdParams p, const __grid_constant__ CUtensorMap tmap, int tiles_per_slab, int n_items) {
    using namespace m256;
    extern __shared__ __align__(128) unsigned char m256_smem[];
    unsigned char* const base = m256_smem + ((128u - (smem_u32(m256_smem) & 127u)) & 127u);
    float2* const stage = reinterpret_cast<float2*>(base);                                      // [P][TR], both halves in turn
    __shared__ float win_sm[kP];
    __shared__ float2 tw_sm[kP];
    __shared__ __align__(8) uint64_t full_bar[2];        // [half]: the half's next tile has landed in the staging buffer
    const int half = threadIdx.x / kHalf;                // warp-uniform
    const int t = threadIdx.x - half * kHalf;
    float2* const sm = reinterpret_cast<float2*>(base + (1 + half) * kTileBytes);               // [P][TR] of this half
    const int rl = t % kTR;
    const int u = t / kTR;
    if (threadIdx.x < 2) mbar_init(&full_bar[threadIdx.x], 1);
    if (threadIdx.x == 0) mbar_fence_init();
    if (threadIdx.x < kP) {
        win_sm[threadIdx.x] = __ldg(p.window + threadIdx.x);
        tw_sm[threadIdx.x] = __ldg(p.tw + (threadIdx.x / kR) * (threadIdx.x % kR));
    }
    __syncthreads();
    // tiles are dealt to the halves in the order half 0 of the CTA, half 1, half 0, ...: use n of the staging buffer belongs to
    // half n & 1 and is its item number n >> 1
    const int worker = (int)blockIdx.x * 2 + half;
    const int stride = (int)gridDim.x * 2;
    const int other = (int)blockIdx.x * 2 + (half ^ 1);
    const MtdHalfSync sync{1 + half};
    if (threadIdx.x == 0) Mtd256Issue{&tmap, stage, &full_bar[0], worker, tiles_per_slab, worker < n_items}();
    const float2* tl = stage + u * kTR + rl;             // pulse u of this thread's column; pulse u + j R is j R rows further on
    int it = 0;
#pragma unroll 1
    for (int item = worker; item < n_items; item += stride, ++it) {
        const int slab = item / tiles_per_slab;
        const int r = (item - slab * tiles_per_slab) * kTR + rl;
        const bool ok = r < p.cols;
        mbar_wait(&full_bar[half], (uint32_t)(it & 1));
        float2 v[kR];
        if (MTI) {
            // x[p + lag] - x[p], the last `lag` pulses are zero (MP/fun_Process_MTI.m:20-22)
            const float2* tl2 = tl + p.mti_lag * kTR;
            const int last = kP - p.mti_lag - u;            // pulse u + j*R is kept iff j*R < last
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                float2 x = make_float2(0.f, 0.f);
                if (j * kR < last) x = csub(tl2[j * kR * kTR], tl[j * kR * kTR]);
                v[j] = x;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kR; ++j) v[j] = tl[j * kR * kTR];
        }
        // the samples are in registers: order these generic-proxy reads before the refill that follows the half's first barrier
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // next user of the staging buffer: the other half's item `it` (half 0) or its item `it + 1` (half 1)
        const int nxt = other + (it + half) * stride;
        const Mtd256Issue hook{&tmap, stage, &full_bar[half ^ 1], nxt, tiles_per_slab, t == 0 && nxt < n_items};
        mtd_fast_body<kR, kTR, CF, true>(v, p, sm, win_sm, tw_sm, slab, r, ok, u, rl, sync, hook);
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
    const int rows_in = p.in_rows > 0 ? p.in_rows : P;
    const float2* col = p.in + (size_t)slab * rows_in * p.in_ld + r;
    for (int prt = threadIdx.y; prt < P; prt += blockDim.y) {
        float2 x = make_float2(0.f, 0.f);
        if (ok && prt < rows_in) x = cscale(mtd_load(col, prt, rows_in, p.in_ld, p.mti_lag), __ldg(p.window + prt));
        a[prt * TR + rl] = x;
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
    if (ok && p.out_c) {
        // slow-time transform FIRST (crop-aware fun_MTD_produce): the caller pulse-compresses only the rows it keeps
        const int half = p.no_shift ? 0 : P / 2;
        const int nrow = p.crop_hi - p.crop_lo + 1;
        float2* out = p.out_c + (size_t)slab * nrow * p.out_ld + r;
        for (int m = threadIdx.y; m < P; m += blockDim.y) {
            int row = m + half;
            if (row >= P) row -= P;
            if (row < p.crop_lo || row > p.crop_hi) continue;
            float2 x = a[m * TR + rl];
            if (row >= p.zv_lo && row <= p.zv_hi) x = make_float2(0.f, 0.f);
            out[(size_t)(row - p.crop_lo) * p.out_ld] = x;
        }
    } else if (ok) {
        float* out = p.out + (size_t)slab * P * p.out_ld + r;
        const int half = p.no_shift ? 0 : P / 2;
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

// Mixed-radix path for the reference's own CPI lengths: P = 16 * 16 * P3 with P3 = 6 (1536 PRTs per frame,
// MP/main_produce_dataset_win_xzr.m:33-38) or P3 = 8 (the DMX script's 2048-point zero-padded transform,
// CW/DMX_SignalProcessing_main_xzr.m:414).  One CTA owns TR = 8 adjacent range cells of one slab and keeps their P x 8
// samples in shared memory; three in-place DIF stages with register butterflies (radix 16, radix 16, radix P3), task =
// (butterfly, column) with the column fastest, so a warp touches four consecutive rows x eight columns = 256 contiguous
// bytes per access.  Twiddle indices need no modulo: stage 1 uses w_P^(u k1) with u k1 < P, stage 2 w_P^(16 v k2).
// Frequency k = k1 + 16 k2 + 256 k3 ends at position k1 * (P/16) + k2 * P3 + k3; the output loop maps it to its row.
// Same options as the generic kernel: zero-padded input rows, no fftshift, MTI, complex cropped output.
template <int P3>
__device__ __forceinline__ void small_dft(float2 (&v)[P3], const float2* __restrict__ w /* w_P3^m, m < P3 */) {
    float2 o[P3];
#pragma unroll
    for (int k = 0; k < P3; ++k) {
        float2 acc = v[0];
#pragma unroll
        for (int j = 1; j < P3; ++j) {
            const float2 t = w[(j * k) % P3];
            acc.x = fmaf(v[j].x, t.x, acc.x); acc.x = fmaf(-v[j].y, t.y, acc.x);
            acc.y = fmaf(v[j].x, t.y, acc.y); acc.y = fmaf(v[j].y, t.x, acc.y);
        }
        o[k] = acc;
    }
#pragma unroll
    for (int k = 0; k < P3; ++k) v[k] = o[k];
}

template <int P3>
__global__ void __launch_bounds__(256, 1) mtd_mixed_kernel(const MtdParams p) {
    constexpr int TR = 8;
    constexpr int P = 256 * P3;
    constexpr int S1 = P / 16;         // stride of the first radix-16 stage
    constexpr int S2 = P3;             // stride of the second
    extern __shared__ float2 sm[];     // [P][TR]
    __shared__ float2 w3[P3];
    const int c = threadIdx.x % TR;
    const int task0 = threadIdx.x / TR;             // 0..31
    constexpr int NT = 256 / TR;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * TR + c;
    const bool ok = r < p.cols;
    const int rows_in = p.in_rows > 0 ? p.in_rows : P;
    if (threadIdx.x < P3) w3[threadIdx.x] = __ldg(p.tw + threadIdx.x * (P / P3));
    {
        const float2* col = p.in + (size_t)slab * rows_in * p.in_ld + (ok ? r : 0);
        for (int prt = task0; prt < P; prt += NT) {
            float2 x = make_float2(0.f, 0.f);
            if (ok && prt < rows_in) x = cscale(mtd_load(col, prt, rows_in, p.in_ld, p.mti_lag), __ldg(p.window + prt));
            sm[prt * TR + c] = x;
        }
    }
    __syncthreads();
    // ---- stage 1: radix 16 over stride S1, twiddle w_P^(u k1)
    for (int u = task0; u < S1; u += NT) {
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = sm[(j * S1 + u) * TR + c];
        Dft<16, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            if (k1 > 0) v[k1] = cmul(v[k1], __ldg(p.tw + u * k1));
            sm[(k1 * S1 + u) * TR + c] = v[k1];
        }
    }
    __syncthreads();
    // ---- stage 2: inside every block k1, radix 16 over stride S2, twiddle w_S1^(v k2) = w_P^(16 v k2)
    for (int t = task0; t < 16 * S2; t += NT) {
        const int k1 = t / S2, vv = t - k1 * S2;
        float2* base = sm + (size_t)(k1 * S1 + vv) * TR + c;
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = base[j * S2 * TR];
        Dft<16, -1>::run(v);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            if (k2 > 0) v[k2] = cmul(v[k2], __ldg(p.tw + 16 * vv * k2));
            base[k2 * S2 * TR] = v[k2];
        }
    }
    __syncthreads();
    // ---- stage 3: radix P3 over the P3 consecutive entries of every block (k1, k2), then the output row
    const int half = p.no_shift ? 0 : P / 2;
    const int nrow = p.crop_hi - p.crop_lo + 1;
    for (int t = task0; t < 256; t += NT) {
        const int k1 = t >> 4, k2 = t & 15;
        float2* base = sm + (size_t)(k1 * S1 + k2 * S2) * TR + c;
        float2 v[P3];
#pragma unroll
        for (int j = 0; j < P3; ++j) v[j] = base[j * TR];
        if (P3 == 8) {
            float2 a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = v[j % P3];
            dft8<-1>(a);
#pragma unroll
            for (int j = 0; j < P3; ++j) v[j] = a[j];
        } else {
            small_dft<P3>(v, w3);
        }
        if (!ok) continue;
#pragma unroll
        for (int k3 = 0; k3 < P3; ++k3) {
            const int k = k1 + 16 * k2 + 256 * k3;
            int row = k + half;
            if (row >= P) row -= P;
            const bool zero = row >= p.zv_lo && row <= p.zv_hi;
            if (p.out_c) {
                if (row < p.crop_lo || row > p.crop_hi) continue;
                p.out_c[((size_t)slab * nrow + (row - p.crop_lo)) * p.out_ld + r] = zero ? make_float2(0.f, 0.f) : v[k3];
            } else {
                p.out[((size_t)slab * P + row) * p.out_ld + r] = zero ? 0.f : hypotf(v[k3].x, v[k3].y);
            }
        }
    }
}

bool mtd_has_fast_path(int P) { return P == 64 || P == 256; }
bool mtd_fast_fuses_cfar(int P) { return P == 64 || P == 256; }
int mtd_generic_max_p() { return 12288; }

template <int R, int TR>
static cudaError_t launch_fast(const MtdParams& p, int n_slabs, cudaStream_t st) {
    constexpr int P = R * R;
    const size_t smem = (size_t)P * TR * sizeof(float2);
    dim3 grid((p.cols + TR - 1) / TR, n_slabs, 1);
    if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
    if ((size_t)P * p.in_ld >= (1ull << 31) || (size_t)P * p.out_ld >= (1ull << 31)) return cudaErrorInvalidValue;
    const int cf = !p.cfar_on ? 0 : (p.cf.ref_v == 5 && p.cf.guard_v == 7) ? 2 : 1;
    if (p.cfar_on && 2 * (p.cf.ref_v + p.cf.guard_v) > P) return cudaErrorInvalidValue;   // window reads stay inside the buffer
#define RB_MTD_FAST(MTI, CF)                                                                                   \
    do {                                                                                                       \
        static size_t configured[64] = {};                                                                     \
        cudaError_t ce = ensure_dynamic_smem(mtd_fast_kernel<R, TR, MTI, CF>, smem, configured);               \
        if (ce != cudaSuccess) return ce;                                                                      \
        mtd_fast_kernel<R, TR, MTI, CF><<<grid, TR * R, smem, st>>>(p);                                        \
    } while (0)
    if (p.mti_lag > 0) {
        if (cf == 0) RB_MTD_FAST(true, 0); else if (cf == 1) RB_MTD_FAST(true, 1); else RB_MTD_FAST(true, 2);
    } else {
        if (cf == 0) RB_MTD_FAST(false, 0); else if (cf == 1) RB_MTD_FAST(false, 1); else RB_MTD_FAST(false, 2);
    }
#undef RB_MTD_FAST
    return cudaGetLastError();
}

// P = 256 through the persistent kernel: needs 16-byte aligned rows for the tensor map
static cudaError_t launch_mtd256_tma(const MtdParams& p, int n_slabs, cudaStream_t st) {
    using namespace m256;
    if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
    if ((size_t)kP * p.in_ld >= (1ull << 31) || (size_t)kP * p.out_ld >= (1ull << 31)) return cudaErrorInvalidValue;
    const int cf = !p.cfar_on ? 0 : (p.cf.ref_v == 5 && p.cf.guard_v == 7) ? 2 : 1;
    if (p.cfar_on && 2 * (p.cf.ref_v + p.cf.guard_v) > kP) return cudaErrorInvalidValue;
    if (p.mti_lag < 0 || p.mti_lag >= kP) return cudaErrorInvalidValue;
    const int tiles_per_slab = (p.cols + kTR - 1) / kTR;
    const long long n_items = (long long)tiles_per_slab * n_slabs;
    if (n_items > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)p.cols * 2, (cuuint64_t)kP, (cuuint64_t)n_slabs};
    const cuuint64_t strides[2] = {(cuuint64_t)p.in_ld * sizeof(float2), (cuuint64_t)kP * p.in_ld * sizeof(float2)};
    const cuuint32_t box[3] = {2 * kTR, (cuuint32_t)kP, 1};
    cudaError_t ce = tensor_map_encode_tiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (ce != cudaSuccess) return ce;
    const int grid = (int)std::min<long long>((n_items + 1) / 2, p.n_sms);     // two tiles in flight per CTA
#define RB_MTD256(MTI, CF)                                                                                     \
    do {                                                                                                       \
        static size_t configured[64] = {};                                                                     \
        ce = ensure_dynamic_smem(mtd256_tma_kernel<MTI, CF>, (size_t)kSmemBytes, configured);                  \
        if (ce != cudaSuccess) return ce;                                                                      \
        mtd256_tma_kernel<MTI, CF><<<grid, kThreads, kSmemBytes, st>>>(p, map, tiles_per_slab, (int)n_items);  \
    } while (0)
    if (p.mti_lag > 0) {
        if (cf == 0) RB_MTD256(true, 0); else if (cf == 1) RB_MTD256(true, 1); else RB_MTD256(true, 2);
    } else {
        if (cf == 0) RB_MTD256(false, 0); else if (cf == 1) RB_MTD256(false, 1); else RB_MTD256(false, 2);
    }
#undef RB_MTD256
    return cudaGetLastError();
}

cudaError_t launch_mtd(const MtdParams& p, int n_slabs, cudaStream_t st) {
    if (p.cols <= 0 || n_slabs <= 0) return cudaSuccess;
    const bool plain = p.in_rows == 0 && !p.no_shift && !p.out_c;
    if (p.P == 64 && plain) return launch_fast<8, 64>(p, n_slabs, st);
    if (p.P == 256 && plain) {
        const bool tma_ok = p.n_sms > 0 && !p.no_tma && (p.in_ld % 2) == 0 && (reinterpret_cast<uintptr_t>(p.in) % 16) == 0;
        return tma_ok ? launch_mtd256_tma(p, n_slabs, st) : launch_fast<16, 32>(p, n_slabs, st);
    }
    if (p.P == 1536 || p.P == 2048) {
        const size_t smem = (size_t)p.P * 8 * sizeof(float2);
        dim3 grid((p.cols + 7) / 8, n_slabs, 1);
        if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
        static size_t configured[2][64] = {};
        if (p.P == 1536) {
            cudaError_t ce = ensure_dynamic_smem(mtd_mixed_kernel<6>, smem, configured[0]);
            if (ce != cudaSuccess) return ce;
            mtd_mixed_kernel<6><<<grid, 256, smem, st>>>(p);
        } else {
            cudaError_t ce = ensure_dynamic_smem(mtd_mixed_kernel<8>, smem, configured[1]);
            if (ce != cudaSuccess) return ce;
            mtd_mixed_kernel<8><<<grid, 256, smem, st>>>(p);
        }
        return cudaGetLastError();
    }
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
