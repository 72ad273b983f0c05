// onepass_kernel.cu -- the single-pass chain for P = 64, 16 interleaved lanes (third version): int16 unpack -> pulse
// compression -> Kaiser window + 64-point slow-time FFT + fftshift + |.| + zero-velocity mask -> velocity CA-CFAR, with the
// pulse-compressed intermediate kept in SHARED MEMORY (it never goes to HBM).
//
// Replaces, in one kernel, the sequence MP/fun_MTD_produce.m:67-79 (fun_lss_pulse_compression -> fun_Process_MTD ->
// fun_0v_pressing) on the unpacked frame of FrameDataRead_xzr.m:138,150-156 and the velocity stage of
// CW/executeCFAR.m:23-31 (CW/Function_CFAR1D_sub.m:17-69).  The sparse range stage (executeCFAR.m:45-89) stays the
// separate kernel cfar_r64_kernel, which consumes the hit list / column masks written here.
//
// Input staging.  The 16 lanes of the wire format are interleaved at 4-byte granularity ([range][lane][I,Q]), so a CTA
// that owns ONE lane cannot fetch its input from the wire buffer efficiently.  A small bandwidth-bound kernel
// (deinterleave_kernel) first rewrites a chunk of CPIs as lane planes  [cpi][lane][prt pair][range (padded)][2 prts]
// of offset-binary int16 pairs (x ^ 0x8000: the later int16 -> fp32 conversion is then two byte permutes and one packed
// subtraction, no I2F); the chunk is small enough that the planes are written and re-read inside L2.  The pad columns hold
// the offset-binary zero, so tiles that run past the PRT read x = 0 (linear, not circular, correlation).
//
// Work decomposition.  An ITEM is (CPI, overlap-save tile, lane): 64 PRT lines x 256 range samples in, 64 Doppler rows x
// V range cells out (V = 256 - taps + 1 rounded down to a multiple of 4, 188 for the 67-tap reference).  One persistent
// CTA per SM (384 threads = 12 warps, 168 registers, ~205 KB of shared memory) walks its items:
//   * the item's 64 KB of input arrives as two TENSOR-MAP TMA loads (cp.async.bulk.tensor.2d, one 256 x 16 box of 8-byte
//     elements each; mbarrier complete_tx), issued one item ahead as soon as the previous contents have been consumed;
//   * PC TASKS: a task is four lines (two PRT pairs).  A warp owns a task, 16 threads per PRT pair, every thread holds the
//     SAME 16 sample positions of TWO lines (radix-16 x 16; forward DIF -> reference spectrum -> inverse DIT), so the
//     twiddles and spectrum values it fetches serve two lines and two independent dependency chains are in flight.  The
//     two exchanges of a transform go through the line's own (padded, 17-slot pitch) slab row with __syncwarp only: there
//     is no CTA-wide barrier anywhere in the kernel.  Constant complex factors are held as two scalars and applied with
//     FMUL2 + FFMA2 (scalar-broadcast and swap/negate operand modifiers), no operand shuffling instructions;
//   * DOPPLER: when the 16 tasks of an item have arrived on an mbarrier, six warps pull one range column per thread out
//     of the slab into registers (64 complex = 128 registers), release the slab through a second mbarrier and run window,
//     64-point FFT, |.|, zero-velocity mask, RDM stores, velocity CFAR, hit word + list append from registers, while the
//     other six warps are already transforming the next item; tasks are handed out by a shared-memory counter and the
//     Doppler warps take tasks again (up to the item they serve next) when their columns are done.
#include "common.cuh"
#include "radix.cuh"
#include "tw64.cuh"
#include "pc_core.cuh"
#include "pcw_core.cuh"
#include "kernels.h"
#include "../../include/radar_b200.h"
#include "tmap.h"
#include <algorithm>

namespace rb {

namespace op {
constexpr int kThreads = 384;
constexpr int kDopplerWarps = 6;          // 6 x 32 columns >= V
constexpr int kP = 64;                    // PRTs per CPI
constexpr int kLanes = 16;                // interleaved lanes of the wire format
constexpr int kNT = 256;                  // overlap-save tile
constexpr int kTasks = 16;                // PC tasks per item: four lines (two PRT pairs) each
constexpr int kRowC = 272;                // complex slots per slab row: the 16 x 16 exchange matrix at a pitch of 17 slots
constexpr int kSlabBytes = kP * kRowC * 8;            // 139 264
constexpr int kRawHalfBytes = 16 * kNT * 8;           // 32 768: 16 PRT pairs x 256 ranges x (2 x int16 pair)
constexpr int kRawBytes = 2 * kRawHalfBytes;          // 65 536
constexpr int kTabBytes = 256 * 8 + 256 * 8;          // twiddles, reference spectrum
constexpr int kSmemBytes = kSlabBytes + kRawBytes + kTabBytes;     // 208 896 (+ static barriers) of 232 448
constexpr int kMaxV = 32 * kDopplerWarps; // 192
}  // namespace op

__device__ __forceinline__ float op_fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// The register-resident Doppler column (one thread = one range cell): v[] holds the 64 windowed slow-time samples.
// 64-point DIF as 8 x radix-8 / compile-time twiddles / 8 x radix-8, |.| with the zero-velocity keep factors, RDM rows
// stored range-contiguous (streaming stores), velocity CA-CFAR on the register column with directly summed windows
// (tested rows 1..63, windows y-12..y-8 and y+8..y+12 with the edge substitution of Function_CFAR1D_sub.m:30-39),
// hit word and detection-list append.  Same arithmetic as mtd64_column (mtd64_core.cuh).
__device__ __forceinline__ void doppler_column(float2 (&v)[64], const OnePassParams& p, int slab_id, int cpi, int member, int r, bool ok,
                                               int lane) {
    constexpr int P = 64, REF = 5, GUARD = 7, N0 = 0;
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
    float mag[P];   // indexed by output row (fftshifted, MP/fun_Process_MTD.m:24)
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) {
        float2 a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = v[q + 8 * k0];
        dft8<-1>(a);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const int row = (k0 + 8 * k1 + P / 2) & (P - 1);
            mag[row] = op_fast_sqrt(a[k1].x * a[k1].x + a[k1].y * a[k1].y) * p.keep[row];   // MP/fun_0v_pressing.m:4-6
        }
    }
    constexpr int NV = P - 2 * N0 - 1;
    // S[i] = mag[i] + ... + mag[i+4] over the cropped axis (summed left to right like the reference's mean); the left window
    // of cell y is S[y-12], the right one S[y+8]: every five-cell sum is formed once and used by two cells
    float S[NV - REF + 1];
#pragma unroll
    for (int i = 0; i + REF <= NV; ++i) {
        float a = mag[N0 + 1 + i];
#pragma unroll
        for (int j = 1; j < REF; ++j) a += mag[N0 + 1 + i + j];
        S[i] = a;
    }
    unsigned long long hits = 0ull;
#pragma unroll
    for (int y = 0; y < NV; ++y) {
        const int l1 = y - GUARD - REF;
        const int r1 = y + GUARD + 1;
        const bool okL = l1 >= 0;
        const bool okR = r1 + REF - 1 <= NV - 1;
        const float sl = okL ? S[okL ? l1 : 0] : 0.f;
        const float sr = okR ? S[okR ? r1 : 0] : 0.f;
        const float a = okL ? sl : sr;
        const float b = okR ? sr : sl;
        const float mu = p.meth_v == 0 ? fmaxf(a, b) : fminf(a, b);
        if (mag[N0 + 1 + y] >= mu * p.tv_over_ref) hits |= 1ull << (N0 + 1 + y);
    }
    {
        int slo, shi;
        if (!ok || !cfar_seg_of(p.segs, r, p.R, &slo, &shi)) hits = 0ull;
    }
    if (ok) p.colmask[(size_t)slab_id * p.R + r] = hits;
    // list slots: one atomic per warp, issued BEFORE the RDM stores so that its round trip hides behind them
    const bool any = __any_sync(0xffffffffu, hits != 0ull);
    const int n = __popcll(hits);
    int incl = n, base = 0;
    if (any) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) base = atomicAdd(p.det_count, incl);
    }
    float* out = p.rdm + (size_t)slab_id * P * p.R + r;
    if (ok) {
#pragma unroll
        for (int row = 0; row < P; ++row) __stcs(out + (size_t)row * p.R, mag[row]);
    }
    if (!any) return;
    base = __shfl_sync(0xffffffffu, base, 31);
    int slot = base + incl - n;
    unsigned long long h = hits;
    while (h) {
        const int row = __ffsll((long long)h) - 1;
        h &= h - 1;
        float amp = 0.f;                               // mag[] is register-resident with static indexing only: select chain
#pragma unroll
        for (int i = 0; i < P; ++i) amp = (i == row) ? mag[i] : amp;
        if (slot < p.max_det) {
            rb200_det d;
            d.cpi = (uint32_t)(p.cpi0 + cpi);
            d.r = (uint32_t)r;
            d.v = (uint16_t)row;
            d.lane = (uint8_t)member;
            d.kind = RB200_DET_V;
            d.amp = amp;
            reinterpret_cast<rb200_det*>(p.dets)[slot] = d;
        }
        ++slot;
    }
}


__global__ void __launch_bounds__(op::kThreads, 1) onepass_kernel(const __grid_constant__ OnePassParams p, const __grid_constant__ CUtensorMap tmap) {
    using namespace op;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* slab = reinterpret_cast<float2*>(smem_raw);                                    // [64][272] complex
    uint2* rawbuf = reinterpret_cast<uint2*>(smem_raw + kSlabBytes);                       // [2 halves][16 PRT pairs][256 ranges] x (prt 2pp, prt 2pp+1)
    float2* tw_sm = reinterpret_cast<float2*>(smem_raw + kSlabBytes + kRawBytes);          // [q][n1]: w256^(n1*q)
    float2* h_sm = tw_sm + 256;                                                            // [j][n1]: spectrum bin n1 + 16*j
    __shared__ __align__(8) uint64_t raw_full[2];      // half h of the item's input has landed (TMA complete_tx)
    __shared__ __align__(8) uint64_t raw_empty[2];     // the 8 tasks of half h have pulled their input into registers
    __shared__ __align__(8) uint64_t slab_full;        // the 16 tasks of the item have stored their compressed lines
    __shared__ __align__(8) uint64_t slab_empty;       // the 6 Doppler warps hold the item's columns in registers
    __shared__ int next_task;                          // monotonic: task g belongs to local item g >> 4

    const int t = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);     // warp-uniform by construction
    const int lane = t & 31;
    const int n1 = lane & 15, hh = lane >> 4;
    const int n_total = p.n_cpi * p.n_tiles * kLanes;
    if ((int)blockIdx.x >= n_total) return;
    const int my_items = (n_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int V = p.V, R = p.R;

    if (t == 0) {
        mbar_init(&raw_full[0], 1);
        mbar_init(&raw_full[1], 1);
        mbar_init(&raw_empty[0], kTasks / 2);
        mbar_init(&raw_empty[1], kTasks / 2);
        mbar_init(&slab_full, kTasks);
        mbar_init(&slab_empty, kDopplerWarps);
        mbar_fence_init();
        next_task = 0;
    }
    if (t < 256) {
        tw_sm[t] = __ldg(p.tw + t);                                           // tw[q*16 + n1]
        h_sm[t] = __ldg(p.hperm + (t & 15) * 16 + (t >> 4));                  // h_sm[j*16 + n1] = bin n1 + 16*j
    }
    __syncthreads();

    // local item k -> (cpi, tile, lane); lanes fastest, so that the 16 lanes of a tile run side by side on 16 SMs
    auto item_of = [&](int k, int& cpi, int& tile, int& ln) {
        const int id = (int)blockIdx.x + k * (int)gridDim.x;
        ln = id & (kLanes - 1);
        const int tg = id >> 4;
        cpi = tg / p.n_tiles;
        tile = tg - cpi * p.n_tiles;
    };
    auto issue_half = [&](int k, int half) {        // one thread
        int cpi, tile, ln;
        item_of(k, cpi, tile, ln);
        mbar_expect_tx(&raw_full[half], (uint32_t)kRawHalfBytes);
        tma_load_2d(reinterpret_cast<unsigned char*>(rawbuf) + half * kRawHalfBytes, &tmap, tile * V, (cpi * kLanes + ln) * 32 + half * 16,
                    &raw_full[half]);
    };
    if (t == 0) {
        issue_half(0, 0);
        issue_half(0, 1);
    }

    const bool is_dop = warp < kDopplerWarps;
    int kD = 0;                                     // Doppler warps: the item whose columns this warp transforms next
#pragma unroll 1
    while (true) {
        // ---- next task: PC warps take them in order without limit, Doppler warps only up to the item they serve next
        int g = -1;
        if (lane == 0) {
            if (!is_dop) {
                g = atomicAdd(&next_task, 1);
            } else {
                const int limit = kTasks * (kD + 1);
                int old = *reinterpret_cast<volatile int*>(&next_task);
                while (old < limit) {
                    const int prev = atomicCAS(&next_task, old, old + 1);
                    if (prev == old) { g = old; break; }
                    old = prev;
                }
            }
        }
        g = __shfl_sync(0xffffffffu, g, 0);

        if (is_dop && g < 0) {
            // ================= Doppler columns of item kD, register-resident =================
            if (kD >= my_items) break;
            int cpi, tile, ln;
            item_of(kD, cpi, tile, ln);
            const int r0 = tile * V;
            const int Vt = min(V, R - r0);
            const int c = 32 * warp + lane;
            const bool ok = c < Vt;
            mbar_wait(&slab_full, (uint32_t)(kD & 1));
            float2 v[64];
            {
                const float2* col = slab + (ok ? c : Vt - 1);
#pragma unroll
                for (int prt = 0; prt < 64; ++prt) v[prt] = cscale(col[prt * kRowC], p.win[prt]);   // MP/fun_Process_MTD.m:22
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&slab_empty);
            if (!(p.dbg & 1)) doppler_column(v, p, cpi * kLanes + ln, cpi, ln, r0 + (ok ? c : Vt - 1), ok, lane);
            ++kD;
            continue;
        }
        if (g >= kTasks * my_items) break;          // PC warps: all tasks handed out

        // ================= PC task g: four lines (PRT pairs 2*tau, 2*tau + 1) of local item k =================
        const int k = g >> 4, tau = g & (kTasks - 1), half = tau >> 3;
        const int ppl = (tau & 7) * 2 + hh;         // PRT pair inside the half
        const int prtA = (half * 16 + ppl) * 2;     // lines prtA, prtA + 1
        if (k > 0) mbar_wait(&slab_empty, (uint32_t)((k - 1) & 1));       // the previous item's columns have left the slab
        mbar_wait(&raw_full[half], (uint32_t)(k & 1));
        float2 a[16], b[16];
        {
            const uint2* rw = rawbuf + half * (kRawHalfBytes / 8) + ppl * kNT + n1;
#pragma unroll
            for (int j = 0; j < 16; ++j) {                                  // x[n1 + 16 j] of both lines (FrameDataRead_xzr.m:154-156)
                const uint2 w = rw[16 * j];
                a[j] = unpack_ob(w.x);
                b[j] = unpack_ob(w.y);
            }
        }
        // this task's input is in registers: when the 8 tasks of a half are, the half is refilled for item k + 1 (generic-
        // proxy reads -> fence -> mbarrier arrive [release]; the warp holding the half's last task waits for the eight
        // arrivals [acquire] and issues the tensor copy [async proxy])
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[half]);
        if ((tau & 7) == 7) {
            mbar_wait(&raw_empty[half], (uint32_t)(k & 1));
            if (k + 1 < my_items && lane == 0) issue_half(k + 1, half);
        }
        if (!(p.dbg & 2)) {
            // exchange rows = the lines' own slab rows (layout in pcw_core.cuh)
            float2* const rowA = slab + prtA * kRowC;
            float2* const rowB = rowA + kRowC;
            PcTwiddles twr;
            twr.load(tw_sm, n1);
            pc_pair_transform(a, b, twr, h_sm, n1, rowA + n1, rowB + n1, rowA + 17 * n1, rowB + 17 * n1);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 12; ++j)
                if (n1 + 16 * j < V) {                                      // alias-free lags only (V <= 192)
                    rowA[n1 + 16 * j] = a[j];
                    rowB[n1 + 16 * j] = b[j];
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&slab_full);
    }
}

// ---- wire -> lane planes.  Thread = (lane quad q, range r) of one PRT pair: two 16-byte loads (4 lanes of PRT 2pp and of
// PRT 2pp + 1; a warp reads 2 x 512 contiguous bytes) and four 8-byte stores (one per lane plane; the eight threads of a
// quad write 64 contiguous bytes).
__global__ void __launch_bounds__(256) deinterleave_kernel(const int4* __restrict__ raw, uint2* __restrict__ planar, int R, int Rp, int n_rblocks) {
    const int t = threadIdx.x;
    const int q = t & 3;
    int bid = blockIdx.x;
    const int rb = bid % n_rblocks;
    bid /= n_rblocks;
    const int pp = bid & 31;
    const int cpi = bid >> 5;
    const int r = rb * 64 + (t >> 2);
    if (r >= R) return;
    const int4 w0 = __ldcs(raw + ((size_t)(cpi * op::kP + 2 * pp) * R + r) * 4 + q);
    const int4 w1 = __ldcs(raw + ((size_t)(cpi * op::kP + 2 * pp + 1) * R + r) * 4 + q);
    uint2* dst = planar + ((size_t)((cpi * op::kLanes + 4 * q) * 32 + pp)) * Rp + r;
    const size_t plane = (size_t)32 * Rp;
    const unsigned X = 0x80008000u;
    dst[0] = make_uint2((unsigned)w0.x ^ X, (unsigned)w1.x ^ X);
    dst[plane] = make_uint2((unsigned)w0.y ^ X, (unsigned)w1.y ^ X);
    dst[2 * plane] = make_uint2((unsigned)w0.z ^ X, (unsigned)w1.z ^ X);
    dst[3 * plane] = make_uint2((unsigned)w0.w ^ X, (unsigned)w1.w ^ X);
}

__global__ void fill_u32_kernel(unsigned* p, size_t n, unsigned v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
int onepass_tile_valid(int n_taps) {
    const int v = (op::kNT - n_taps + 1) & ~3;
    return std::min(v, op::kMaxV);
}

int onepass_planar_pitch(int R, int V) {            // 8-byte elements per (lane, PRT pair) row: every tile's 256-sample window fits
    const int n_tiles = (R + V - 1) / V;
    return ((n_tiles - 1) * V + op::kNT + 1) & ~1;
}

size_t onepass_planar_bytes(int n_cpi, int R, int V) { return (size_t)n_cpi * op::kLanes * 32 * onepass_planar_pitch(R, V) * 8; }

cudaError_t onepass_planar_init(void* planar, size_t bytes, cudaStream_t st) {
    fill_u32_kernel<<<1024, 256, 0, st>>>(reinterpret_cast<unsigned*>(planar), bytes / 4, 0x80008000u);
    return cudaGetLastError();
}

cudaError_t launch_deinterleave(const void* raw, void* planar, int n_cpi, int R, int V, cudaStream_t st) {
    const int n_rblocks = (R + 63) / 64;
    deinterleave_kernel<<<(unsigned)(n_cpi * 32 * n_rblocks), 256, 0, st>>>(reinterpret_cast<const int4*>(raw), reinterpret_cast<uint2*>(planar), R,
                                                                            onepass_planar_pitch(R, V), n_rblocks);
    return cudaGetLastError();
}

// tensor map of the lane planes: 2-D, 8-byte elements, [rows = cpi*16*32 + lane*32 + prt pair][pitch]; box = 256 x 16
static cudaError_t encode_planar_map(CUtensorMap* map, void* planar, int n_cpi, int R, int V) {
    const int pitch = onepass_planar_pitch(R, V);
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)n_cpi * op::kLanes * 32};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 8};
    const cuuint32_t box[2] = {(cuuint32_t)op::kNT, 16};
    return tensor_map_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, planar, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

cudaError_t launch_onepass(const OnePassParams& p, void* planar, int n_sms, cudaStream_t st) {
    static size_t configured[64] = {};
    cudaError_t ce = ensure_dynamic_smem(onepass_kernel, (size_t)op::kSmemBytes, configured);
    if (ce != cudaSuccess) return ce;
    if (p.n_cpi < 1 || p.n_tiles < 1 || p.V < 4 || p.V > op::kMaxV || (p.V & 3) || (p.R & 3)) return cudaErrorInvalidValue;
    alignas(64) CUtensorMap map;
    ce = encode_planar_map(&map, planar, p.n_cpi, p.R, p.V);
    if (ce != cudaSuccess) return ce;
    const int n_items = p.n_cpi * p.n_tiles * op::kLanes;
    const int grid = std::max(1, std::min(n_sms, n_items));
    onepass_kernel<<<grid, op::kThreads, op::kSmemBytes, st>>>(p, map);
    return cudaGetLastError();
}

}  // namespace rb
