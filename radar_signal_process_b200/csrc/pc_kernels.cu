// pc_kernels.cu -- K1: int16 I/Q unpack (+ iSTC gain) + segmented pulse compression.
//
// Replaces the per-PRT loop of MP/fun_lss_pulse_compression.m:24-37 (9-arg twin
// MTD/fun_lss_pulse_compression.m:36-65) and the three per-PRT FFTs of
// MP/fun_pulse_compression.m:19-22, fused with the DDC unpack of FrameDataRead_xzr.m:138,150-156
// and the optional iSTC gain of MP/fun_iSTC.m:14.
//
// Algorithm: overlap-save fast correlation.  Every waveform segment is cut into tiles of NT = R^S
// input samples; one CTA transforms LT independent lines (the 16 interleaved channels of one PRT in
// wire mode) of one tile:
//     in-place radix-R DIF forward FFT (S stages; stage 0 straight from global memory with the
//     int16 -> fp32 unpack, later stages exchanged through shared memory)
//   -> multiply by the resident, digit-reversed reference spectrum conj(FFT(taps)) * scale/NT
//   -> in-place radix-R DIT inverse FFT back to natural order
//   -> the V = NT-L+1 alias-free lags are stored range-contiguous as float2 [line][range].
// The forward transform leaves the spectrum digit-reversed and the inverse consumes it that way, so
// no reordering pass exists; 2(S-1) shared-memory exchanges per tile.
//
// Thread mapping: lane-fastest (thread = lane + LT*butterfly) for every stage but the last, so that
// a warp's global load covers 2 range cells x 16 channels = 128 contiguous bytes of the wire format
// and shared-memory accesses of a half-warp hit 16 different lanes (odd line stride -> no bank
// conflicts).  The final inverse stage is re-mapped butterfly-fastest so each warp stores 256
// contiguous bytes of one output line.
#include "common.cuh"
#include "radix.cuh"
#include "pc_core.cuh"
#include "kernels.h"
#include <algorithm>
#include <cmath>
#include <vector>

namespace rb {

template <int R, int S, int LT, bool WIRE, int CFIX>
__global__ void __launch_bounds__(LT * (ipow(R, S) / R), PcOcc<R, S, LT>::min_blocks)
pc_fft_kernel(const PcParams p) {
    constexpr int NT = ipow(R, S);
    constexpr int NB = NT / R;     // butterflies per line per stage
    extern __shared__ float2 sm[];

    const int t = threadIdx.x;
    const int lane = t % LT;
    const int u = t / LT;
    const int2 tile = __ldg(&p.tiles[blockIdx.y]);
    const PcSegDev& sg = p.segs[tile.x];
    const int g = blockIdx.x;
    const int in_off = tile.y * sg.V - sg.pre;
    const int C = CFIX > 0 ? CFIX : p.C;

    float2 v[R];
    // ---- stage 0 operands straight from global memory (unpack fused) ----
    {
        const int lane_g = blockIdx.z * LT + lane;   // wire: channel index
        bool lane_ok;
        long long e0;        // index of the element (range = in_start + in_off + u) of this line
        int es;              // element stride between consecutive range cells
        if (WIRE) {
            lane_ok = lane_g < C;
            e0 = ((long long)g * p.R + sg.in_start + in_off + u) * C + lane_g;
            es = C;
        } else {
            const int line = g * LT + lane;
            lane_ok = line < p.n_lines;
            e0 = (long long)line * p.R + sg.in_start + in_off + u;
            es = 1;
        }
        const bool interior = in_off >= 0 && in_off + NT <= sg.in_len;   // CTA-uniform
        if (interior && lane_ok) {
            if (WIRE) {
                const int* src = reinterpret_cast<const int*>(p.in) + e0;
                int w[R];
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = __ldg(src + (long long)j * NB * es);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    v[j].x = (float)(short)(w[j] & 0xffff);   // I (FrameDataRead_xzr.m:154)
                    v[j].y = (float)(w[j] >> 16);             // Q (:155)
                }
            } else {
                const float2* src = reinterpret_cast<const float2*>(p.in) + e0;
#pragma unroll
                for (int j = 0; j < R; ++j) v[j] = __ldg(src + j * NB);
            }
        } else {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int rs = in_off + u + j * NB;
                float2 x = make_float2(0.f, 0.f);
                if (lane_ok && rs >= 0 && rs < sg.in_len) {
                    if (WIRE) {
                        const int w = __ldg(reinterpret_cast<const int*>(p.in) + e0 + (long long)j * NB * es);
                        x.x = (float)(short)(w & 0xffff);
                        x.y = (float)(w >> 16);
                    } else {
                        x = __ldg(reinterpret_cast<const float2*>(p.in) + e0 + j * NB);
                    }
                }
                v[j] = x;
            }
        }
        if (p.gain) {      // iSTC (MP/fun_iSTC.m:14): per-range gain before compression
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int rs = in_off + u + j * NB;
                if (rs >= 0 && rs < sg.in_len) v[j] = cscale(v[j], __ldg(p.gain + sg.in_start + rs));
            }
        }
    }
    int out_lane, out_u;
    pc_fft_core<R, S, LT, false>(v, sm, p.tw, p.hperm + sg.h_off, t, lane, u, out_lane, out_u);

    // ---- store the alias-free lags ----
    {
        size_t oline;
        bool ok;
        if (WIRE) {
            const int cpi = g / p.P, prt = g % p.P;
            const int out_lane_g = blockIdx.z * LT + out_lane;
            ok = out_lane_g < C;
            oline = ((size_t)cpi * C + out_lane_g) * p.P + prt;
        } else {
            const int line = g * LT + out_lane;
            ok = line < p.n_lines;
            oline = line;
        }
        if (ok) {
            const int n0 = tile.y * sg.V;
            float2* o = p.out + oline * p.R_out + sg.out_start;
            if (sg.rot == 0) {
                const int lim = min(sg.V, sg.out_len - n0);    // valid lags of this tile
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
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant for the wire format with 16 interleaved channels: the raw tile of a work item
// (NT range cells x 16 channels x 4 B, contiguous in HBM) is fetched by one TMA bulk copy
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) into a staging buffer.  A tile is consumed
// into registers at the very top of its item, so the SAME buffer is refilled for the next item as soon as
// every thread has arrived on the `empty` barrier -- the copy then has the whole transform of the current
// item to land, and one 16 KB buffer is enough.  With 53 KB of shared memory and 64 registers per thread
// four CTAs (32 warps) are resident per SM.  Grid = resident CTAs; items (line group, tile) round-robin.
// ---------------------------------------------------------------------------------------------
#ifndef RB_PC_TMA_MINB
#define RB_PC_TMA_MINB 4
#endif
template <int R, int S, bool GAIN>
__global__ void __launch_bounds__(16 * (ipow(R, S) / R), RB_PC_TMA_MINB)
pc_fft_tma_kernel(const PcParams p, int n_items, int n_tiles, int h_entries) {
    constexpr int LT = 16;
    constexpr int NT = ipow(R, S);
    constexpr int NB = NT / R;
    constexpr int LS = NT + 1;
    constexpr int FFT_BYTES = ((LT * LS * (int)sizeof(float2)) + 127) / 128 * 128;
    constexpr int RAW_INTS = NT * LT;                 // one staged tile: [range][channel] int16 pairs
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* sm = reinterpret_cast<float2*>(smem_raw);
    int* rawbuf = reinterpret_cast<int*>(smem_raw + FFT_BYTES);
    // resident tables: stage-major twiddles and up to TAB_SEGS segment spectra (each NT float2)
    constexpr int TW_N = tw_block_off<R, S>(S - 1) > 0 ? tw_block_off<R, S>(S - 1) : 1;
    float2* tw_sm = reinterpret_cast<float2*>(smem_raw + FFT_BYTES + RAW_INTS * 4);
    float2* h_sm = tw_sm + TW_N;
    __shared__ __align__(8) uint64_t mbar[2];          // [0] full (TMA bytes landed), [1] empty (every thread has read the tile)
    __shared__ float gain_sm[GAIN ? 2 * NT : 1];       // iSTC gains of the current / next tile (blockDim.x == NT threads fill it)

    const int t = threadIdx.x;
    const int lane = t % LT;
    const int u = t / LT;
    if (t == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], (int)blockDim.x);
        mbar_fence_init();
    }
    for (int i = t; i < TW_N; i += blockDim.x) tw_sm[i] = __ldg(p.tw + i);
    for (int i = t; i < h_entries; i += blockDim.x) h_sm[i] = __ldg(p.hperm + i);
    // gains of tile sample t of `item` (MP/fun_iSTC.m:14); samples outside the segment get 0 (their raw value is 0 too)
    auto stage_gain = [&](int item, int buf) {
        const int g = item / n_tiles;
        const int2 tile = __ldg(&p.tiles[item - g * n_tiles]);
        const PcSegDev& sg = p.segs[tile.x];
        const int rs = tile.y * sg.V - sg.pre + t;
        gain_sm[GAIN ? buf * NT + t : 0] = (rs >= 0 && rs < sg.in_len) ? __ldg(p.gain + sg.in_start + rs) : 0.f;
    };
    if (GAIN && (int)blockIdx.x < n_items) stage_gain(blockIdx.x, 0);
    __syncthreads();

    // thread 0: start the bulk copy of an item's valid input span into the staging buffer
    auto issue = [&](int item) {
        const int g = item / n_tiles;
        const int2 tile = __ldg(&p.tiles[item - g * n_tiles]);
        const PcSegDev& sg = p.segs[tile.x];
        const int in_off = tile.y * sg.V - sg.pre;
        const int lo = max(in_off, 0);
        const int hi = min(in_off + NT, sg.in_len);
        if (hi > lo) {
            const uint32_t bytes = (uint32_t)(hi - lo) * (LT * 4);
            mbar_expect_tx(&mbar[0], bytes);
            const int* src = reinterpret_cast<const int*>(p.in) + ((size_t)g * p.R + sg.in_start + lo) * LT;
            bulk_g2s(rawbuf + (lo - in_off) * LT, src, bytes, &mbar[0]);
        } else {
            mbar_arrive(&mbar[0]);
        }
    };

    int item = blockIdx.x;
    if (t == 0 && item < n_items) issue(item);
    for (int it = 0; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const int g = item / n_tiles;
        const int2 tile = __ldg(&p.tiles[item - g * n_tiles]);
        const PcSegDev& sg = p.segs[tile.x];
        const int in_off = tile.y * sg.V - sg.pre;

        mbar_wait(&mbar[0], (uint32_t)(it & 1));
        float2 v[R];
        {
            const int* rb = rawbuf + u * LT + lane;
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
            if (GAIN) {      // staged one item ahead (below), so no global-load latency sits in front of the first butterfly
                const float* gs = gain_sm + (GAIN ? buf * NT + u : 0);
#pragma unroll
                for (int j = 0; j < R; ++j) v[j] = cscale(v[j], gs[GAIN ? j * NB : 0]);
            }
        }
        // generic-proxy reads of the staging buffer are ordered before the TMA (async-proxy) refill: every thread fences
        // and arrives on `empty`; thread 0 refills the buffer for the next item once all arrivals are in
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&mbar[1]);
        {
            const int next = item + gridDim.x;
            if (t == 0) {
                mbar_wait(&mbar[1], (uint32_t)(it & 1));
                if (next < n_items) issue(next);
            }
            if (GAIN && next < n_items) stage_gain(next, buf ^ 1);      // read after the end-of-item barrier
        }
        int out_lane, out_u;
        pc_fft_core<R, S, LT, true>(v, sm, tw_sm, h_sm + sg.h_off, t, lane, u, out_lane, out_u);
        {
            const int cpi = g / p.P, prt = g - cpi * p.P;
            const size_t oline = ((size_t)cpi * LT + out_lane) * p.P + prt;
            const int n0 = tile.y * sg.V;
            float2* o = p.out + oline * p.R_out + sg.out_start;
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
        __syncthreads();     // the exchange buffer is reused by the next item
    }
}

// Time-domain fallback for tap counts beyond the largest FFT tile: one thread per output sample.
template <bool WIRE>
__global__ void pc_direct_kernel(const PcParams p, const float2* __restrict__ taps, int seg_idx, int n_groups_lines) {
    const PcSegDev& sg = p.segs[seg_idx];
    const int n = blockIdx.y * blockDim.x + threadIdx.x;
    const int line = blockIdx.x;   // wire: group*C + lane ; planar: line
    if (n >= sg.out_len) return;
    size_t ibase, oline;
    int stride;
    if (WIRE) {
        const int g = line / p.C, lane = line % p.C;
        ibase = (size_t)g * p.R * p.C + lane;
        stride = p.C;
        const int cpi = g / p.P, prt = g % p.P;
        oline = ((size_t)cpi * p.C + lane) * p.P + prt;
    } else {
        ibase = (size_t)line * p.R;
        stride = 1;
        oline = line;
    }
    float2 acc = make_float2(0.f, 0.f);
    const float2* tp = taps + sg.t_off;
    for (int k = 0; k < sg.n_taps; ++k) {
        const int rs = n + k - sg.pre;
        if (rs < 0 || rs >= sg.in_len) continue;
        const int r = sg.in_start + rs;
        float2 x;
        if (WIRE) {
            const int w = __ldg(reinterpret_cast<const int*>(p.in) + ibase + (size_t)r * stride);
            x = make_float2((float)(short)(w & 0xffff), (float)(short)(w >> 16));
        } else {
            x = __ldg(reinterpret_cast<const float2*>(p.in) + ibase + r);
        }
        if (p.gain) x = cscale(x, __ldg(p.gain + r));
        const float2 c = cmulc(x, __ldg(tp + k));
        acc.x += c.x;
        acc.y += c.y;
    }
    int c = n - sg.rot;
    if (c < 0) c += sg.out_len;
    p.out[oline * p.R_out + sg.out_start + c] = acc;
}

// zero the output columns no segment writes (s_PC_0 = zeros(...), MP/fun_lss_pulse_compression.m:18)
__global__ void pc_zero_cols_kernel(float2* out, size_t n_lines, int R, int c0, int c1) {
    const int w = c1 - c0;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines * (size_t)w) return;
    const size_t line = i / w;
    const int c = c0 + (int)(i % w);
    out[line * R + c] = make_float2(0.f, 0.f);
}

// plain unpack (parity/debug entry rb200_unpack_ddc_i16): wire -> float2 [cpi][lane][prt][range]
__global__ void unpack_kernel(const int* __restrict__ raw, float2* __restrict__ out, int n_groups, int P, int R, int C) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into wire words
    const size_t total = (size_t)n_groups * R * C;
    if (i >= total) return;
    const int lane = (int)(i % C);
    const size_t gr = i / C;
    const int r = (int)(gr % R);
    const size_t g = gr / R;
    const int cpi = (int)(g / P), prt = (int)(g % P);
    const int w = __ldg(raw + i);
    out[(((size_t)cpi * C + lane) * P + prt) * R + r] = make_float2((float)(short)(w & 0xffff), (float)(short)(w >> 16));
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int R, int S, int LT, bool WIRE, int CFIX>
static cudaError_t launch_fft(const PcParams& p, int n_tiles, int n_groups, cudaStream_t st) {
    constexpr int NT = ipow(R, S);
    constexpr int threads = LT * (NT / R);
    const size_t smem = (size_t)LT * (NT + 1) * sizeof(float2);
    static size_t configured[64] = {};   // per template instantiation and device
    cudaError_t ce = ensure_dynamic_smem(pc_fft_kernel<R, S, LT, WIRE, CFIX>, smem, configured);
    if (ce != cudaSuccess) return ce;
    dim3 grid(n_groups, n_tiles, WIRE ? (p.C + LT - 1) / LT : 1);
    if (n_tiles > 65535) return cudaErrorInvalidConfiguration;
    pc_fft_kernel<R, S, LT, WIRE, CFIX><<<grid, threads, smem, st>>>(p);
    return cudaGetLastError();
}

// Host: stage-major twiddle table consumed by pc_fft_kernel (see tw_block_off).
void pc_build_twiddles(int nt, std::vector<float2>& tw) {
    const int R = (nt == 512) ? 8 : 16;
    int S = 0;
    for (int n = nt; n > 1; n /= R) ++S;
    tw.clear();
    int Rs = 1;                                   // R^s
    for (int s = 0; s + 1 < S; ++s) {
        const int st = nt / (Rs * R);
        for (int k = 0; k < R; ++k)
            for (int q = 0; q < st; ++q) {
                const double a = -2.0 * M_PI * (double)(((long long)q * k * Rs) % nt) / (double)nt;
                tw.push_back(make_float2((float)cos(a), (float)sin(a)));
            }
        Rs *= R;
    }
    if (tw.empty()) tw.push_back(make_float2(1.f, 0.f));
}

cudaError_t launch_pc_fft_tma(const PcParams& p, int n_tiles, int n_groups, int n_sms, int ctas_per_sm, int h_entries, cudaStream_t st) {
    constexpr int R = 16, S = 2, NT = 256, LT = 16;
    const size_t fft_bytes = ((size_t)LT * (NT + 1) * sizeof(float2) + 127) / 128 * 128;
    const size_t smem = fft_bytes + (size_t)NT * LT * 4 + ((size_t)tw_block_off<R, S>(S - 1) + h_entries) * sizeof(float2);
    const long long n_items = (long long)n_tiles * n_groups;
    if (n_items <= 0 || n_items > 0x7fffffffLL) return n_items <= 0 ? cudaSuccess : cudaErrorInvalidConfiguration;
    // resident CTAs per SM as the hardware will actually place them (shared memory grows with the number of segment spectra)
    auto resident = [&](auto kernel, size_t (&configured)[64]) -> int {
        if (ensure_dynamic_smem(kernel, smem, configured) != cudaSuccess) return -1;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, LT * (NT / R), smem) != cudaSuccess) return -1;
        return std::max(1, std::min(std::min(ctas_per_sm, RB_PC_TMA_MINB), nb));
    };
    if (p.gain) {
        static size_t configured[64] = {};
        const int per_sm = resident(pc_fft_tma_kernel<R, S, true>, configured);
        if (per_sm < 0) return cudaGetLastError() != cudaSuccess ? cudaErrorInvalidValue : cudaErrorInvalidValue;
        const int grid = (int)std::min<long long>(n_items, (long long)n_sms * per_sm);
        pc_fft_tma_kernel<R, S, true><<<grid, LT * (NT / R), smem, st>>>(p, (int)n_items, n_tiles, h_entries);
    } else {
        static size_t configured[64] = {};
        const int per_sm = resident(pc_fft_tma_kernel<R, S, false>, configured);
        if (per_sm < 0) return cudaErrorInvalidValue;
        const int grid = (int)std::min<long long>(n_items, (long long)n_sms * per_sm);
        pc_fft_tma_kernel<R, S, false><<<grid, LT * (NT / R), smem, st>>>(p, (int)n_items, n_tiles, h_entries);
    }
    return cudaGetLastError();
}

int pc_tile_lanes(int nt, bool wire) {
    (void)wire;
    switch (nt) {
        case 256: return 16;
        case 512: return 16;
        case 4096: return 2;
        default: return 0;
    }
}

cudaError_t launch_pc_fft(int nt, bool wire, const PcParams& p, int n_tiles, int n_groups, cudaStream_t st) {
    if (nt == 256) {
        if (!wire) return launch_fft<16, 2, 16, false, 0>(p, n_tiles, n_groups, st);
        return p.C == 16 ? launch_fft<16, 2, 16, true, 16>(p, n_tiles, n_groups, st) : launch_fft<16, 2, 16, true, 0>(p, n_tiles, n_groups, st);
    }
    if (nt == 512) {
        if (!wire) return launch_fft<8, 3, 16, false, 0>(p, n_tiles, n_groups, st);
        return p.C == 16 ? launch_fft<8, 3, 16, true, 16>(p, n_tiles, n_groups, st) : launch_fft<8, 3, 16, true, 0>(p, n_tiles, n_groups, st);
    }
    if (nt == 4096) return wire ? launch_fft<16, 3, 2, true, 0>(p, n_tiles, n_groups, st) : launch_fft<16, 3, 2, false, 0>(p, n_tiles, n_groups, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_pc_direct(bool wire, const PcParams& p, const float2* taps, int seg_idx, int out_len, int n_lines, cudaStream_t st) {
    dim3 grid(n_lines, (out_len + 127) / 128, 1);
    if (grid.y > 65535) return cudaErrorInvalidConfiguration;
    if (wire) pc_direct_kernel<true><<<grid, 128, 0, st>>>(p, taps, seg_idx, n_lines);
    else pc_direct_kernel<false><<<grid, 128, 0, st>>>(p, taps, seg_idx, n_lines);
    return cudaGetLastError();
}

cudaError_t launch_pc_zero_cols(float2* out, size_t n_lines, int R, int c0, int c1, cudaStream_t st) {
    if (c1 <= c0 || n_lines == 0) return cudaSuccess;
    const size_t total = n_lines * (size_t)(c1 - c0);
    pc_zero_cols_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, n_lines, R, c0, c1);
    return cudaGetLastError();
}

cudaError_t launch_unpack(const int16_t* raw, float2* out, int n_groups, int P, int R, int C, cudaStream_t st) {
    const size_t total = (size_t)n_groups * R * C;
    unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const int*>(raw), out, n_groups, P, R, C);
    return cudaGetLastError();
}

}  // namespace rb
