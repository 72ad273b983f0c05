// chain64_kernel.cu -- one persistent kernel for a whole batch of P = 64, 16-channel CPIs: pulse compression
// and the MTD + velocity-CFAR stage run as two roles of the same resident CTAs, with the pulse-compressed
// intermediate kept in a three-CPI ring that lives in the 126 MB L2 instead of making a round trip through HBM.
//
// Work items are taken from one global counter in an order that never makes a CTA wait in steady state:
//     PC(0), PC(1), MTD(0), PC(2), MTD(1), ..., PC(B-1), MTD(B-2), MTD(B-1)
// so every MTD(c) item is fetched a full PC phase after the last PC(c) item was handed out, and PC(c) overwrites
// ring slot c mod 3 a full phase after MTD(c-3) finished reading it.  Completion counters per CPI (release:
// __threadfence + atomicAdd, acquire: volatile spin + __threadfence) make the dependency explicit; the spins
// are bounded and raise the context error flag instead of hanging.  All CTAs are resident (grid = SMs x 2), so
// the waits cannot deadlock.  Intermediate reads use ld.global.cg (L1 is not coherent across SMs).
//
//   role PC  : pc_fft_core on one (PRT, 256-sample tile, 16 channels) item, raw tile prefetched with a TMA bulk
//              copy into a double-buffered staging area (same as pc_fft_tma_kernel)
//   role MTD : mtd64_column on 256 range columns of one (CPI, lane) slab (same as mtd64_kernel)
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "mtd64_core.cuh"
#include "pc_core.cuh"

namespace rb {

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// bounded wait until *ctr >= want (thread 0 only)
__device__ __forceinline__ void wait_counter(const int* ctr, int want, int* err_flag) {
    int spins = 0;
    while (ld_volatile(ctr) < want) {
        __nanosleep(64);
        if (++spins > (1 << 22)) {       // ~0.3 s: something is badly wrong; do not hang the GPU
            *err_flag = 2;
            break;
        }
    }
    __threadfence();
}

template <int R, int S>
__global__ void __launch_bounds__(256, 2)
chain64_kernel(const Chain64Params q) {
    constexpr int LT = 16;
    constexpr int NT = ipow(R, S);
    constexpr int NB = NT / R;
    constexpr int LS = NT + 1;
    constexpr int P = 64;
    constexpr int FFT_BYTES = ((LT * LS * (int)sizeof(float2)) + 127) / 128 * 128;
    constexpr int RAW_INTS = NT * LT;
    static_assert(LT * NB == 256, "256 threads per CTA");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    int* rawbuf = reinterpret_cast<int*>(smem_raw + FFT_BYTES);
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ int s_next, s_is_pc, s_cpi, s_local;

    const PcParams& p = q.pc;
    const int t = threadIdx.x;
    const int lane = t % LT;
    const int u = t / LT;
    const int nPC = q.n_pc_items, nMTD = q.n_mtd_items, B = q.n_cpi;
    const int T = nPC + nMTD;
    const int total = B * T;                  // < 2^31 (checked by the launcher)

    if (t == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // item index -> (is_pc, cpi, local index)
    auto decode = [&](int it, bool& is_pc, int& cpi, int& local) {
        if (it < nPC) { is_pc = true; cpi = 0; local = it; return; }
        const int r = it - nPC;
        int blk = r / T;
        if (blk > B - 1) blk = B - 1;
        const int off = r - blk * T;
        if (blk < B - 1 && off < nPC) { is_pc = true; cpi = blk + 1; local = off; }
        else { is_pc = false; cpi = blk; local = (blk < B - 1) ? off - nPC : off; }
    };
    // thread 0: claim the next item and, if it is a PC item, start the TMA copy of its raw tile (reads only the
    // immutable input, so it needs no dependency).  Dependencies are awaited by wait_deps() only after the
    // current item has been completed and signalled, which keeps the scheme deadlock-free.
    int next_buf = 0;              // thread 0 only: staging buffer of the next PC item
    auto claim = [&]() -> int {
        const int it = atomicAdd(q.work_counter, 1);
        if (it >= total) return total;
        bool is_pc; int cpi, local;
        decode(it, is_pc, cpi, local);
        if (is_pc) {
            const int g = local / q.n_tiles;
            const int2 tile = __ldg(&p.tiles[local - g * q.n_tiles]);
            const PcSegDev& sg = p.segs[tile.x];
            const int in_off = tile.y * sg.V - sg.pre;
            const int lo = max(in_off, 0);
            const int hi = min(in_off + NT, sg.in_len);
            const int buf = next_buf;
            next_buf ^= 1;
            if (hi > lo) {
                const uint32_t bytes = (uint32_t)(hi - lo) * (LT * 4);
                mbar_expect_tx(&mbar[buf], bytes);
                const int* src = reinterpret_cast<const int*>(p.in) + (((size_t)cpi * P + g) * p.R + sg.in_start + lo) * LT;
                bulk_g2s(rawbuf + buf * RAW_INTS + (lo - in_off) * LT, src, bytes, &mbar[buf]);
            } else {
                mbar_arrive(&mbar[buf]);
            }
        }
        return it;
    };
    // thread 0 only.  Completion is monotonic, so what is known to be complete is remembered and re-polled
    // only when a new CPI is reached (once per CPI per CTA instead of once per item).
    int known_pc = -1, known_mtd = -1;            // highest CPI index known complete for each role
    auto wait_deps = [&](int it) {
        if (it >= total) return;
        bool is_pc; int cpi, local;
        decode(it, is_pc, cpi, local);
        if (is_pc) {
            const int need = cpi - q.ring_slots;  // ring slot free again?
            if (need > known_mtd) { wait_counter(q.mtd_done + need, nMTD, q.err_flag); known_mtd = need; }
        } else {
            if (cpi > known_pc) { wait_counter(q.pc_done + cpi, nPC, q.err_flag); known_pc = cpi; }   // all lines of this CPI compressed
        }
        s_is_pc = is_pc; s_cpi = cpi; s_local = local;
    };

    // every thread tracks the staging-buffer sequence of the PC items this CTA executes
    int my_buf = 0;
    int my_uses[2] = {0, 0};
    int item;
    if (t == 0) {
        item = claim();
        wait_deps(item);
        s_next = item;
    }
    __syncthreads();
    item = s_next;
    while (item < total) {
        const bool is_pc = s_is_pc != 0;
        const int cpi = s_cpi, local = s_local;
        int nxt = 0;
        if (is_pc) {
            const int g = local / q.n_tiles;
            const int2 tile = __ldg(&p.tiles[local - g * q.n_tiles]);
            const PcSegDev& sg = p.segs[tile.x];
            const int in_off = tile.y * sg.V - sg.pre;
            const int buf = my_buf;
            my_buf ^= 1;
            mbar_wait(&mbar[buf], (uint32_t)(my_uses[buf] & 1));
            ++my_uses[buf];
            float2 v[R];
            {
                const int* rb = rawbuf + buf * RAW_INTS + u * LT + lane;
                const bool interior = in_off >= 0 && in_off + NT <= sg.in_len;
                if (interior) {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int w = rb[j * NB * LT];
                        v[j].x = (float)(short)(w & 0xffff);   // I (FrameDataRead_xzr.m:154)
                        v[j].y = (float)(w >> 16);             // Q (:155)
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int rs = in_off + u + j * NB;
                        int w = 0;
                        if (rs >= 0 && rs < sg.in_len) w = rb[j * NB * LT];
                        v[j].x = (float)(short)(w & 0xffff);
                        v[j].y = (float)(w >> 16);
                    }
                }
                if (p.gain) {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int rs = in_off + u + j * NB;
                        if (rs >= 0 && rs < sg.in_len) v[j] = cscale(v[j], __ldg(p.gain + sg.in_start + rs));
                    }
                }
            }
            // generic-proxy reads of the staging buffer are ordered before the TMA (async-proxy) refill that thread 0
            // issues for a later item
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (t == 0) nxt = claim();            // prefetch: the other staging buffer was consumed one PC item ago
            int out_lane, out_u;
            pc_fft_core<R, S, LT, false>(v, sm, p.tw, p.hperm + sg.h_off, t, lane, u, out_lane, out_u);
            {
                float2* ring = q.ring + (size_t)(cpi % q.ring_slots) * q.ring_stride;
                const size_t oline = (size_t)out_lane * P + g;             // [lane][prt] inside the ring slot
                const int n0 = tile.y * sg.V;
                float2* o = ring + oline * p.R_out + sg.out_start;
                if (sg.rot == 0) {
                    const int lim = min(sg.V, sg.out_len - n0);
                    o += n0 + out_u;
#pragma unroll
                    for (int j = 0; j < R; ++j)
                        if (out_u + j * NB < lim) o[j * NB] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int nl = out_u + j * NB;
                        const int n = n0 + nl;
                        if (nl < sg.V && n < sg.out_len) {
                            int c = n - sg.rot;
                            if (c < 0) c += sg.out_len;
                            o[c] = v[j];
                        }
                    }
                }
            }
            __threadfence();                      // release this thread's stores before the completion count
        } else {
            // ---- MTD role: 256 range columns of slab (cpi, lane_idx) ----
            const int tiles_per_slab = (p.R + 255) / 256;
            const int lane_idx = local / tiles_per_slab;
            const int r = (local - lane_idx * tiles_per_slab) * 256 + t;
            const bool ok = r < p.R;
            const int rc = ok ? r : p.R - 1;
            if (t == 0) nxt = claim();
            const float2* col = q.ring + (size_t)(cpi % q.ring_slots) * q.ring_stride + (size_t)lane_idx * P * p.R_out + rc;
            float2 v[P];
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int prt = qq + 8 * j;
                    const float2 x = __ldcg(col + (size_t)prt * p.R_out);
                    v[prt] = make_float2(x.x * q.m.win[prt], x.y * q.m.win[prt]);
                }
            mtd64_column<5, 7, 0, true>(v, q.m, cpi * LT + lane_idx, r, ok);
            __threadfence();
        }
        __syncthreads();                          // exchange buffer / staging reuse; all stores of the item issued
        if (t == 0) {
            atomicAdd(is_pc ? q.pc_done + cpi : q.mtd_done + cpi, 1);     // this item is complete and visible
            wait_deps(nxt);                                               // only now may this CTA wait on others
            s_next = nxt;
        }
        __syncthreads();
        item = s_next;
    }
}

cudaError_t launch_chain64(const Chain64Params& q, int n_sms, cudaStream_t st) {
    constexpr int R = 16, S = 2, NT = 256, LT = 16;
    const size_t fft_bytes = ((size_t)LT * (NT + 1) * sizeof(float2) + 127) / 128 * 128;
    const size_t smem = fft_bytes + 2 * (size_t)NT * LT * 4;
    static size_t configured[64] = {};
    cudaError_t ce = ensure_dynamic_smem(chain64_kernel<R, S>, smem, configured);
    if (ce != cudaSuccess) return ce;
    const long long total = (long long)q.n_cpi * (q.n_pc_items + q.n_mtd_items);
    if (total <= 0) return cudaSuccess;
    if (total > 0x3fffffffLL) return cudaErrorInvalidConfiguration;
    const int grid = (int)std::min<long long>(total, (long long)n_sms * 2);
    chain64_kernel<R, S><<<grid, 256, smem, st>>>(q);
    return cudaGetLastError();
}

}  // namespace rb
