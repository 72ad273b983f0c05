// mtd64_kernel.cu -- K2 for P = 64: Kaiser window + 64-point slow-time FFT + fftshift + |.| + zero-velocity
// mask, fused with the velocity-axis CA-CFAR stage and its detection compaction.
//
// Replaces the per-range-cell loop of MP/fun_Process_MTD.m:20-30, MP/fun_0v_pressing.m:4-6 and the
// velocity stage of CW/executeCFAR.m:28 (CW/Function_CFAR1D_sub.m:17-69).
//
// One thread owns one range cell of one slab and keeps its whole 64-sample Doppler column in
// registers (64 complex = 128 registers; B200 has 64 K registers per SM):
//   * 64 loads of 8 B, each warp-coalesced along range (256 B per warp per PRT row);
//   * window weights and the zero-velocity keep factors come from the kernel parameter block
//     (constant bank operands, no loads);
//   * the 64-point FFT is 8 x radix-8, compile-time twiddles (immediates), 8 x radix-8 -- no shared
//     memory, no barriers;
//   * fftshift is output indexing; magnitudes are stored range-contiguous (128 B per warp per row);
//   * the velocity CFAR runs on the magnitudes still in registers: a register prefix sum gives every
//     window sum with two subtractions; edge substitution and the tested-row crop are resolved at
//     compile time for the (ref, guard, n0) specialisation;
//   * per column the 64 hit bits are stored as one 8-byte word (consumed by the range stage for
//     de-duplication) and hits are appended to the detection list with one atomic per warp.
#include "common.cuh"
#include "radix.cuh"
#include "tw64.cuh"
#include "mtd64_core.cuh"
#include "kernels.h"
#include "../../include/radar_b200.h"
#include <algorithm>

namespace rb {

template <int REF, int GUARD, int N0, bool CFAR>
__global__ void __launch_bounds__(128, RB200_MTD64_MINB)
mtd64_kernel(const Mtd64Params p) {
    constexpr int P = 64;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * 128 + threadIdx.x;
    const bool ok = r < p.cols;
    const int rc = ok ? r : p.cols - 1;      // clamp: inactive threads still take part in warp votes
    const float2* col = p.in + (size_t)slab * P * p.in_ld + rc;
    float2 v[P];
    // loads in the order the first butterflies consume them (q-major: rows q, q+8, ..., q+56)
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int prt = q + 8 * j;
            const float2 x = __ldg(col + (size_t)prt * p.in_ld);
            v[prt] = make_float2(x.x * p.win[prt], x.y * p.win[prt]);
        }
    mtd64_column<REF, GUARD, N0, CFAR>(v, p, slab, r, ok);
}

// Persistent variant: the 64 x 128 tile of the next work item is fetched by 64 TMA bulk copies (one 1 KB row
// each, one mbarrier) into shared memory while the CTA transforms the current item out of registers, so the
// HBM/L2 read stream overlaps the butterflies.  Needs 16-byte aligned rows (even in_ld and cols).
__device__ __forceinline__ uint32_t m64_smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

template <int REF, int GUARD, int N0, bool CFAR>
__global__ void __launch_bounds__(128, RB200_MTD64_MINB)
mtd64_tma_kernel(const Mtd64Params p, int tiles_per_slab, int n_items) {
    constexpr int P = 64;
    extern __shared__ __align__(128) float2 tile[];    // [64][128]
    // full : completes when the 64 row copies of an item have landed (TMA complete_tx)
    // empty: completes when all 128 threads have pulled their column of the current item into registers; the
    //        producers wait on it before overwriting the tile (generic-proxy reads -> async-proxy writes need a
    //        real acquire: a bare bar.sync let a late warp observe rows of the next item)
    __shared__ __align__(8) uint64_t full_bar, empty_bar;
    const int t = threadIdx.x;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m64_smem_u32(&full_bar)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(m64_smem_u32(&empty_bar)), "r"(128));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait_bar = [&](uint64_t* bar, uint32_t parity) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}" ::"r"(m64_smem_u32(bar)),
            "r"(parity)
            : "memory");
    };
    auto expect = [&](int item) {          // thread 0
        const int c0 = (item % tiles_per_slab) * 128;
        const uint32_t bytes = (uint32_t)min(128, p.cols - c0) * 8u * P;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m64_smem_u32(&full_bar)), "r"(bytes) : "memory");
    };
    auto copy_row = [&](int item) {        // threads 0..63: one PRT row each
        const int slab = item / tiles_per_slab;
        const int c0 = (item - slab * tiles_per_slab) * 128;
        const uint32_t bytes = (uint32_t)min(128, p.cols - c0) * 8u;
        const float2* src = p.in + ((size_t)slab * P + t) * p.in_ld + c0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(m64_smem_u32(tile + t * 128)),
                     "l"(src), "r"(bytes), "r"(m64_smem_u32(&full_bar))
                     : "memory");
    };
    // RB200_SPLIT: the producer kernel (pcw_kernel on other SMs) counts finished warp tasks per CPI; a copy thread fetches its
    // row only when the item's CPI is complete (acquire load, bounded spin), then orders the peers' generic-proxy stores
    // before its own async-proxy read
    auto ready = [&](int item) {
        if (!p.wait_done) return;
        const int cpi = (item / tiles_per_slab) / p.n_lanes;
        const int* flag = p.wait_done + cpi;
        int spins = 0, v;
        while (true) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v >= p.wait_target) break;
            if (++spins > (1 << 24)) {
                if (p.err_flag) atomicExch(p.err_flag, 2);
                break;
            }
            __nanosleep(200);
        }
        asm volatile("fence.proxy.async;" ::: "memory");
    };
    int item = blockIdx.x;
    if (item < n_items) {
        if (t == 0) expect(item);
        if (t < P) {
            ready(item);
            copy_row(item);
        }
    }
    for (int it = 0; item < n_items; item += gridDim.x, ++it) {
        const int slab = item / tiles_per_slab;
        const int r = (item - slab * tiles_per_slab) * 128 + t;
        const bool ok = r < p.cols;
        wait_bar(&full_bar, (uint32_t)(it & 1));
        float2 v[P];
#pragma unroll
        for (int prt = 0; prt < P; ++prt) {
            const float2 x = tile[prt * 128 + t];
            v[prt] = make_float2(x.x * p.win[prt], x.y * p.win[prt]);
        }
        // this thread's column is in registers: order the generic-proxy reads above before the async-proxy (TMA) writes
        // that the producers will issue once everybody has arrived, then release the tile
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(m64_smem_u32(&empty_bar)) : "memory");
        const int next = item + gridDim.x;
        if (next < n_items && t < P) {
            if (t == 0) expect(next);
            wait_bar(&empty_bar, (uint32_t)(it & 1));      // every thread of the CTA is done with the tile
            ready(next);
            copy_row(next);
        }
        mtd64_column<REF, GUARD, N0, CFAR>(v, p, slab, r, ok);
    }
}

// REF > 0: the reference-cell count as a compile-time constant, so that both window loops unroll and their loads are
// issued back to back (one memory round trip per decision instead of one per reference cell -- the sparse range stage
// is latency-bound); REF = 0: run-time `ref`.
template <int REF>
__device__ __forceinline__ bool cfar_decide_f32(const float* __restrict__ row, int y, int N, int ref_rt, int guard, float thr, int method, int* err_flag) {
    const int ref = REF > 0 ? REF : ref_rt;
    const int l1 = y - guard - ref;
    const int r1 = y + guard + 1;
    const bool okL = l1 >= 0;
    const bool okR = (y + guard + ref) <= N - 1;
    if (!okL && !okR) {
        if (err_flag) *err_flag = 1;
        return false;
    }
    float sl = 0.f, sr = 0.f;
    const float x = __ldg(row + y);
    if (REF > 0) {
        float wl[REF > 0 ? REF : 1], wr[REF > 0 ? REF : 1];
#pragma unroll
        for (int j = 0; j < REF; ++j) {
            wl[j] = okL ? __ldg(row + l1 + j) : 0.f;
            wr[j] = okR ? __ldg(row + r1 + j) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < REF; ++j) {
            sl += wl[j];
            sr += wr[j];
        }
    } else {
        if (okL) for (int j = 0; j < ref; ++j) sl += __ldg(row + l1 + j);
        if (okR) for (int j = 0; j < ref; ++j) sr += __ldg(row + r1 + j);
    }
    const float mr = sr / (float)ref, ml = sl / (float)ref;
    const float a = okL ? ml : mr;
    const float b = okR ? mr : ml;
    const float mu = method == 0 ? fmaxf(a, b) : fminf(a, b);
    return x >= mu * thr;
}

// row points at the first cell of the range segment, r is relative to it, N is the segment length
template <int REF>
__device__ __forceinline__ int cfar_elect_f32(const float* __restrict__ row, int r, int N, const CfarParams& p, float t_r, int* err_flag) {
    // the three decisions are evaluated unconditionally first (their loads overlap), then combined in column order
    bool pass[3];
    float x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int c = r + d - 1;
        const bool inside = c >= 0 && c < N;
        pass[d] = inside && cfar_decide_f32<REF>(row, inside ? c : r, N, p.ref_r, p.guard_r, t_r, p.meth_r, err_flag);
        x[d] = inside ? __ldg(row + c) : 0.f;
    }
    int best = -1;
    float bestv = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        if (pass[d] && (best < 0 || x[d] > bestv)) { best = r + d - 1; bestv = x[d]; }
    return best;
}

// One list slot for every thread of the warp that is currently converged here, with a single atomic per warp
// (same-address atomics serialise in L2: thousands of hits per chunk would otherwise queue up one by one).
__device__ __forceinline__ int warp_agg_slot(int* counter) {
    const unsigned am = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(am) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(am));
    base = __shfl_sync(am, base, leader);
    return base + __popc(am & ((1u << lane) - 1u));
}

// Range stage for the fused path (CW/executeCFAR.m:45-75): one (grid-stride) thread per velocity hit of
// this chunk; elects the first maximum among the passing cells of {r-1,r,r+1}, de-duplicates through the
// per-column velocity-hit masks, appends the 2-D record to the global list and moves the velocity record
// from the per-slot scratch list to the global one.  The last block to finish re-arms the slot counters.
template <int REF>
__global__ void cfar_r64_kernel(const float* __restrict__ rdm, const CfarParams p, float t_r,
                                const rb200_det* __restrict__ slot_v, int* __restrict__ slot_count,
                                rb200_det* __restrict__ dets_v, rb200_det* __restrict__ dets_2d, int* __restrict__ gcount,
                                const unsigned long long* __restrict__ colmask, int cols_ld, int* err_flag) {
    int n = slot_count[0];
    const int true_n = n;
    if (n > p.max_det) n = p.max_det;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        rb200_det h = slot_v[i];
        if (!p.range_stage) h.kind = RB200_DET_V | RB200_DET_2D;   // executeCFAR.m:91: the final matrix IS the velocity matrix
        {
            const int slot = warp_agg_slot(&gcount[0]);
            if (slot < p.max_det) dets_v[slot] = h;
        }
        if (!p.range_stage) continue;
        const int slab = (int)(h.cpi - p.cpi0) * p.n_lanes + h.lane;
        const int v = h.v, r = (int)h.r;
        int slo, shi;
        if (!cfar_seg_of(p.segs, r, p.R, &slo, &shi)) continue;          // cannot happen: K2 drops such hits
        const float* row = rdm + ((size_t)slab * p.V + v) * p.R + slo;       // the hit's range segment
        const int N = shi - slo;
        const int crel = cfar_elect_f32<REF>(row, r - slo, N, p, t_r, err_flag);
        if (crel < 0) continue;
        const int c = crel + slo;
        const unsigned long long* cm = colmask + (size_t)slab * cols_ld;
        bool owner = true;
        for (int rr = c - 1; rr < r && owner; ++rr) {
            if (rr < slo) continue;
            if (!((cm[rr] >> v) & 1ull)) continue;
            if (cfar_elect_f32<REF>(row, rr - slo, N, p, t_r, nullptr) == crel) owner = false;
        }
        if (!owner) continue;
        const int slot = warp_agg_slot(&gcount[1]);
        if (slot < p.max_det) {
            rb200_det d;
            d.cpi = h.cpi;
            d.r = (uint32_t)c;
            d.v = h.v;
            d.lane = h.lane;
            d.kind = RB200_DET_2D;
            d.amp = row[crel];
            dets_2d[slot] = d;
        }
    }
    // hits dropped by the per-slot capacity still count (the host reports RB200_ERR_OVERFLOW)
    if (blockIdx.x == 0 && threadIdx.x == 0 && true_n > n) atomicAdd(&gcount[0], true_n - n);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&slot_count[1], 1);
        if (ticket == (int)gridDim.x - 1) {      // last block: re-arm the slot for its next chunk
            slot_count[0] = 0;
            slot_count[1] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
bool mtd64_fused_supported(int P, int ref_v, int guard_v, int n0, int mti_lag) {
    return P == 64 && ref_v == 5 && guard_v == 7 && n0 == 0 && mti_lag == 0;
}

cudaError_t launch_mtd64(const Mtd64Params& p, int n_slabs, bool with_cfar, cudaStream_t st) {
    if (p.cols <= 0 || n_slabs <= 0) return cudaSuccess;
    dim3 grid((p.cols + 127) / 128, n_slabs, 1);
    if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
    if (with_cfar) mtd64_kernel<5, 7, 0, true><<<grid, 128, 0, st>>>(p);
    else mtd64_kernel<5, 7, 0, false><<<grid, 128, 0, st>>>(p);
    return cudaGetLastError();
}

// RB200_SPLIT: holds a stream until the producer kernel of a chunk is running (its CTAs are then resident, so the consumer
// grid launched next on this stream can only take the SMs the producer left free)
__global__ void wait_flag_kernel(const int* flag, int* err_flag) {
    int spins = 0, v;
    while (true) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v != 0) break;
        if (++spins > (1 << 24)) {
            if (err_flag) atomicExch(err_flag, 2);
            break;
        }
        __nanosleep(500);
    }
}
cudaError_t launch_wait_flag(const int* flag, int* err_flag, cudaStream_t st) {
    wait_flag_kernel<<<1, 1, 0, st>>>(flag, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_mtd64_tma(const Mtd64Params& p, int n_slabs, int n_sms, int ctas_per_sm, cudaStream_t st) {
    if (p.cols <= 0 || n_slabs <= 0) return cudaSuccess;
    const int tiles_per_slab = (p.cols + 127) / 128;
    const long long n_items = (long long)tiles_per_slab * n_slabs;
    if (n_items > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const size_t smem = 64 * 128 * sizeof(float2);
    static size_t configured[64] = {};
    static bool carve[64] = {};         // function attributes are per device
    cudaError_t ce = ensure_dynamic_smem(mtd64_tma_kernel<5, 7, 0, true>, smem, configured);
    if (ce != cudaSuccess) return ce;
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!carve[dev]) {  // lets a CTA of this kernel join an SM that pcw_shared_kernel configured for the maximum carve-out
        cudaFuncSetAttribute(mtd64_tma_kernel<5, 7, 0, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carve[dev] = true;
    }
    const int per_sm = std::max(1, std::min(ctas_per_sm, RB200_MTD64_MINB));
    const int grid = (int)std::min<long long>(n_items, (long long)n_sms * per_sm);
    mtd64_tma_kernel<5, 7, 0, true><<<grid, 128, smem, st>>>(p, tiles_per_slab, (int)n_items);
    return cudaGetLastError();
}

cudaError_t launch_cfar_r64(const float* rdm, const CfarParams& p, float t_r, const void* slot_v, int* slot_count, void* dets_v,
                            void* dets_2d, int* gcount, const unsigned long long* colmask, int cols_ld, int* err_flag, int n_blocks,
                            cudaStream_t st) {
    if (p.ref_r == 5)
        cfar_r64_kernel<5><<<n_blocks, 128, 0, st>>>(rdm, p, t_r, (const rb200_det*)slot_v, slot_count, (rb200_det*)dets_v,
                                                     (rb200_det*)dets_2d, gcount, colmask, cols_ld, err_flag);
    else
        cfar_r64_kernel<0><<<n_blocks, 128, 0, st>>>(rdm, p, t_r, (const rb200_det*)slot_v, slot_count, (rb200_det*)dets_v,
                                                     (rb200_det*)dets_2d, gcount, colmask, cols_ld, err_flag);
    return cudaGetLastError();
}

}  // namespace rb
