// pcw_kernel.cu -- K1, warp-private variant: int16 unpack (+ iSTC) + segmented overlap-save pulse compression for the wire
// format with 16 interleaved lanes and 256-sample tiles.
//
// Same arithmetic and the same segment semantics as pc_fft_tma_kernel (pc_kernels.cu; MP/fun_lss_pulse_compression.m:24-37,
// MP/fun_pulse_compression.m:4-22, FrameDataRead_xzr.m:150-156, MP/fun_iSTC.m:14), different mapping onto the SM:
//   * one persistent CTA per SM, 16 warps, no CTA-wide barrier after start-up.  Four warps share a 16 KB staging slot that
//     receives the raw tile of a work item (256 range cells x 16 lanes) through a TENSOR-MAP TMA load (3-D map over
//     [line group][range / 2][2 x 16 lanes], 128-byte swizzle; cells before / after the PRT are zero-filled by the copy
//     engine) and is refilled as soon as the four warps have pulled their samples into registers;
//   * a warp owns a lane QUAD of the tile: thread (n1, pair) holds the 16 sample positions n1 + 16 j of TWO lines (lanes
//     4Q + 2 pair, + 1), so twiddles and spectrum values are fetched once for two lines and two independent dependency
//     chains are in flight.  Even n1 sit in the lower half-warp, odd n1 in the upper one: with the 128-byte swizzle every
//     64-bit read of the staging slot and every access of the exchange rows is bank-conflict free;
//   * the two exchanges of a transform go through rows private to the 16 threads of a line pair (__syncwarp only), the
//     int16 -> fp32 conversion is PRMT/LOP3 + one packed subtraction (no I2F), constant factors are scalar pairs applied
//     with FMUL2 + FFMA2, all shared-memory addresses are base + immediate (pcw_core.cuh).
#include "common.cuh"
#include "radix.cuh"
#include "pc_core.cuh"
#include "pcw_core.cuh"
#include "kernels.h"
#include "tmap.h"
#include <algorithm>

namespace rb {

namespace pcw {
constexpr int kWarpsFull = 16;            // alone on the SM: 16 warps, 128 registers, 4 staging slots
constexpr int kWarpsShared = 12;          // next to one mtd64_tma CTA (RB200_COEXIST): 12 warps, 112 registers, 3 staging slots
constexpr int kNT = 256;
constexpr int kLanes = 16;
constexpr int kTileBytes = kNT * kLanes * 4;          // 16 384
constexpr int kExSlots = 3 * kPcwRowC + 1 + kPcwRowC + 1;    // per warp: rows A0, B0, (one slot of skew) A1, B1 -> 1 090 slots
constexpr int kExBytes = kExSlots * 8;                // 8 720
constexpr int kMaxH = 2048;                           // resident spectrum entries (8 segments)
}  // namespace pcw

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

template <bool GAIN, int kWarps>
__device__ __forceinline__ void pcw_body(const PcParams& p, const CUtensorMap& tmap, int n_items, int n_tiles, int h_entries) {
    using namespace pcw;
    constexpr int kSlots = kWarps / 4;        // staging slots; warps 4s .. 4s+3 work on slot s
    constexpr int kThreads = 32 * kWarps;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the 128-byte swizzle pattern repeats every 1024 bytes of SHARED address: align the slots explicitly
    unsigned char* const smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* tiles = smem_al;                                                          // [4 slots][16 KB], swizzled rows of 128 B
    float2* exch = reinterpret_cast<float2*>(smem_al + kSlots * kTileBytes);                 // [16 warps][1090]
    float2* tw_sm = reinterpret_cast<float2*>(smem_al + kSlots * kTileBytes + kWarps * kExBytes);    // [q][n1]: w256^(n1*q)
    float2* h_sm = tw_sm + 256;                                                              // per segment: [j][n1] = bin n1 + 16*j
    __shared__ __align__(8) uint64_t full_bar[kSlots];     // the slot's tile has landed (TMA complete_tx)
    __shared__ __align__(8) uint64_t empty_bar[kSlots];    // the slot's four warps hold their samples in registers

    const int t = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
    const int lane = t & 31;
    const int n1 = ((lane & 7) << 1) | (lane >> 4);        // lower half-warp: even positions, upper: odd
    const int pair = (lane >> 3) & 1;
    const int slot = warp >> 2, quad = warp & 3;
    const int c_lane = quad * 4 + pair * 2;                // the thread's lines: wire lanes c_lane, c_lane + 1

    if (t < kSlots) {
        mbar_init(&full_bar[t], 1);
        mbar_init(&empty_bar[t], 4);
    }
    if (t == 0) mbar_fence_init();
    for (int i = t; i < 256; i += kThreads) tw_sm[i] = __ldg(p.tw + i);
    for (int i = t; i < h_entries; i += kThreads) h_sm[(i & ~255) + (i & 15) * 16 + ((i >> 4) & 15)] = __ldg(p.hperm + i);
    __syncthreads();

    if (p.started && blockIdx.x == 0 && t == 0) {
        *reinterpret_cast<volatile int*>(p.started) = 1;
        __threadfence();
    }
    const int worker = (int)blockIdx.x * kSlots + slot;
    const int stride = (int)gridDim.x * kSlots;
    unsigned char* const tile_sm = tiles + slot * kTileBytes;

    auto issue = [&](int item) {                  // one thread
        const int g = item / n_tiles;
        const int2 tile = __ldg(&p.tiles[item - g * n_tiles]);
        const PcSegDev& sg = p.segs[tile.x];
        const int r_first = sg.in_start + tile.y * sg.V - sg.pre;      // even by construction (launch_pcw checks the plan)
        mbar_expect_tx(&full_bar[slot], (uint32_t)kTileBytes);
        tma_load_3d(tile_sm, &tmap, 0, r_first >> 1, g, &full_bar[slot]);
    };
    if (quad == 0 && lane == 0 && worker < n_items) issue(worker);

    // per-thread constants of the staging read: row (n1 >> 1) + 8 j, 16-byte chunk ((n1 & 1) * 4 + quad) ^ (row & 7), 8 bytes per pair
    const uint2* const rd = reinterpret_cast<const uint2*>(tile_sm + (n1 >> 1) * 128 + ((((n1 & 1) << 2) | quad) ^ (n1 >> 1)) * 16 + pair * 8);
    float2* const ex = exch + warp * kExSlots + pair * (2 * kPcwRowC + 1);
    float2* const rowA = ex;
    float2* const rowB = ex + kPcwRowC;

    PcTwiddles twr;                               // loaded once: the thread keeps its positions n1 + 16 j for every item
    twr.load(tw_sm, n1);
    int it = 0;
#pragma unroll 1
    for (int item = worker; item < n_items; item += stride, ++it) {
        const int g = item / n_tiles;
        const int2 tile = __ldg(&p.tiles[item - g * n_tiles]);
        const PcSegDev& sg = p.segs[tile.x];
        const int in_off = tile.y * sg.V - sg.pre;
        mbar_wait(&full_bar[slot], (uint32_t)(it & 1));
        float2 a[16], b[16];
        // iSTC gains (MP/fun_iSTC.m:14) are per range cell, shared by every PRT and lane: they stay L1-resident, so they are
        // fetched right where they are used instead of being held in registers across the wait
        const float* gp = GAIN ? p.gain + sg.in_start + in_off + n1 : nullptr;
        if (in_off >= 0 && in_off + kNT <= sg.in_len) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {        // x[n1 + 16 j] of both lines (FrameDataRead_xzr.m:154-156)
                const uint2 w = rd[128 * j];
                a[j] = unpack_tc(w.x);
                b[j] = unpack_tc(w.y);
                if (GAIN) {
                    const float gj = __ldg(gp + 16 * j);
                    a[j] = cscale(a[j], gj);
                    b[j] = cscale(b[j], gj);
                }
            }
        } else {                                  // samples outside the segment are zero (the tile may touch its neighbours)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int rs = in_off + n1 + 16 * j;
                const bool in = rs >= 0 && rs < sg.in_len;
                uint2 w = make_uint2(0u, 0u);
                if (in) w = rd[128 * j];
                a[j] = unpack_tc(w.x);
                b[j] = unpack_tc(w.y);
                if (GAIN) {
                    const float gj = in ? __ldg(gp + 16 * j) : 0.f;
                    a[j] = cscale(a[j], gj);
                    b[j] = cscale(b[j], gj);
                }
            }
        }
        // generic-proxy reads of the slot are ordered before its refill by the copy engine: fence, arrive [release]; the
        // slot's first warp waits for the four arrivals [acquire] and issues the next tensor copy [async proxy]
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
        if (quad == 0) {
            mbar_wait(&empty_bar[slot], (uint32_t)(it & 1));
            if (item + stride < n_items && lane == 0) issue(item + stride);
        }
        pc_pair_transform(a, b, twr, h_sm + sg.h_off, n1, rowA + n1, rowB + n1, rowA + 17 * n1, rowB + 17 * n1);
        __syncwarp();                             // the rows are reused by the next item
        {
            const int cpi = g / p.P, prt = g - cpi * p.P;
            const size_t oline = ((size_t)cpi * kLanes + c_lane) * p.P + prt;
            const int n0 = tile.y * sg.V;
            float2* oa = p.out + oline * p.R_out + sg.out_start;
            const size_t ob = (size_t)p.P * p.R_out;                   // next lane
            if (sg.rot == 0) {
                const int lim = min(sg.V, sg.out_len - n0);
                oa += n0 + n1;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n1 + 16 * j < lim) {
                        oa[16 * j] = a[j];
                        oa[16 * j + ob] = b[j];
                    }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int nl = n1 + 16 * j;
                    const int n = n0 + nl;
                    if (nl < sg.V && n < sg.out_len) {
                        int c = n - sg.rot;
                        if (c < 0) c += sg.out_len;
                        oa[c] = a[j];
                        oa[c + ob] = b[j];
                    }
                }
            }
            if (p.cpi_done) {                     // consumer kernels on other SMs wait for this count (RB200_SPLIT)
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(p.cpi_done + cpi, 1);
                }
            }
        }
    }
}

template <bool GAIN>
__global__ void __launch_bounds__(32 * pcw::kWarpsFull, 1)
pcw_kernel(const __grid_constant__ PcParams p, const __grid_constant__ CUtensorMap tmap, int n_items, int n_tiles, int h_entries) {
    pcw_body<GAIN, pcw::kWarpsFull>(p, tmap, n_items, n_tiles, h_entries);
}
// 12 warps x 112 registers + 159 KB of shared memory leave room for one 128-thread mtd64_tma CTA (166 registers, 64 KB) on the
// same SM: the fp32-bound transform of chunk i+1 then runs beside the HBM-bound Doppler kernel of chunk i
template <bool GAIN>
__global__ void __maxnreg__(112)
pcw_shared_kernel(const __grid_constant__ PcParams p, const __grid_constant__ CUtensorMap tmap, int n_items, int n_tiles, int h_entries) {
    pcw_body<GAIN, pcw::kWarpsShared>(p, tmap, n_items, n_tiles, h_entries);
}

// ---------------------------------------------------------------------------------------------
// host side
bool pcw_plan_supported(const PcParams& p, int n_segs, int h_entries) {
    if (p.C != pcw::kLanes || (p.R & 1) || h_entries > pcw::kMaxH || (h_entries & 255)) return false;
    for (int i = 0; i < n_segs; ++i) {
        const PcSegDev& s = p.segs[i];
        if (s.out_len == 0) continue;
        if ((s.V & 1) || ((s.in_start - s.pre) & 1) || (s.h_off & 255)) return false;
    }
    return true;
}

static cudaError_t encode_wire_map(CUtensorMap* map, const void* in, int R, int n_groups) {
    // [group][range / 2][32 words]: a row is two range cells x 16 lanes x (I, Q) = 128 bytes
    const cuuint64_t dims[3] = {32, (cuuint64_t)(R / 2), (cuuint64_t)n_groups};
    const cuuint64_t strides[2] = {128, (cuuint64_t)R * 64};
    const cuuint32_t box[3] = {32, (cuuint32_t)(pcw::kNT / 2), 1};
    return tensor_map_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

cudaError_t launch_pcw(const PcParams& p, int n_tiles, int n_groups, int n_sms, int h_entries, bool shared_sm, cudaStream_t st) {
    const int warps = shared_sm ? pcw::kWarpsShared : pcw::kWarpsFull;
    const int slots = warps / 4;
    const size_t smem = 1024 + (size_t)slots * pcw::kTileBytes + (size_t)warps * pcw::kExBytes + 256 * sizeof(float2) +
                        (size_t)h_entries * sizeof(float2);
    static size_t configured[4][64] = {};
    cudaError_t ce;
    if (shared_sm) {
        // the SM must be configured for the maximum shared-memory carve-out while this kernel runs, or the Doppler CTA that is
        // meant to join it (64 KB) cannot be placed until the SM drains
        static bool carve[2][64] = {};            // function attributes are per device
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (!carve[p.gain ? 1 : 0][dev]) {
            if (p.gain) cudaFuncSetAttribute(pcw_shared_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            else cudaFuncSetAttribute(pcw_shared_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carve[p.gain ? 1 : 0][dev] = true;
        }
        ce = p.gain ? ensure_dynamic_smem(pcw_shared_kernel<true>, smem, configured[3]) : ensure_dynamic_smem(pcw_shared_kernel<false>, smem, configured[2]);
    }
    else ce = p.gain ? ensure_dynamic_smem(pcw_kernel<true>, smem, configured[1]) : ensure_dynamic_smem(pcw_kernel<false>, smem, configured[0]);
    if (ce != cudaSuccess) return ce;
    alignas(64) CUtensorMap map;
    ce = encode_wire_map(&map, p.in, p.R, n_groups);
    if (ce != cudaSuccess) return ce;
    const int n_items = n_tiles * n_groups;
    const int grid = std::max(1, std::min(n_sms, (n_items + slots - 1) / slots));
    if (shared_sm) {
        if (p.gain) pcw_shared_kernel<true><<<grid, 32 * warps, smem, st>>>(p, map, n_items, n_tiles, h_entries);
        else pcw_shared_kernel<false><<<grid, 32 * warps, smem, st>>>(p, map, n_items, n_tiles, h_entries);
    } else {
        if (p.gain) pcw_kernel<true><<<grid, 32 * warps, smem, st>>>(p, map, n_items, n_tiles, h_entries);
        else pcw_kernel<false><<<grid, 32 * warps, smem, st>>>(p, map, n_items, n_tiles, h_entries);
    }
    return cudaGetLastError();
}

}  // namespace rb
