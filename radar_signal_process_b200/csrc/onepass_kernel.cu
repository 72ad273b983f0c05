// onepass_kernel.cu -- the single-pass chain for P = 64, 16 interleaved lanes: int16 unpack -> pulse compression ->
// Kaiser window + 64-point slow-time FFT + fftshift + |.| + zero-velocity mask -> velocity CA-CFAR, with the
// pulse-compressed intermediate kept in SHARED MEMORY (it never goes to HBM).
//
// Replaces, in one kernel, the sequence MP/fun_MTD_produce.m:67-79 (fun_lss_pulse_compression -> fun_Process_MTD ->
// fun_0v_pressing) on the unpacked frame of FrameDataRead_xzr.m:138,150-156 and the velocity stage of
// CW/executeCFAR.m:23-31 (CW/Function_CFAR1D_sub.m:17-69).  The sparse range stage (executeCFAR.m:45-89) stays the
// separate kernel cfar_r64_kernel, which consumes the hit list / column masks written here.
//
// Work decomposition.  An ITEM is (CPI, lane, overlap-save tile): 64 PRT lines x 256 range samples of one lane in,
// 64 Doppler rows x V range cells out (V = 256 - taps + 1 rounded down to a multiple of 4, 188 for the 67-tap
// reference).  One CTA (512 threads, 1 per SM, ~225 KB of shared memory) owns an item:
//   phase 1  pulse compression of the 64 lines, WARP-PRIVATE: a warp transforms two lines at a time, 16 threads x 16
//            points per line (radix-16 x 16, forward DIF -> reference spectrum -> inverse DIT); the two exchanges of a
//            transform go through the line's own row of the slab (in place, XOR-swizzled, __syncwarp only), so the
//            sixteen warps of a CTA drift through different phases and the fp32 and shared-memory pipes overlap
//            instead of alternating.  The thread's 15 twiddles and 16 spectrum values are register-resident per item.
//            int16 -> fp32 uses the 2^23 magic-number construction (PRMT/LOP3 + one packed FADD2; no I2F).
//   phase 2  the slab [64 PRT][V] complex is read column-wise; a PAIR of threads owns a Doppler column (even / odd
//            PRTs, a 32-point register FFT each, one 16-value shuffle exchange, radix-2 combine), magnitudes go back
//            to the slab rows as dense fp32 rows, which (a) leave for HBM as 64 bulk stores (cp.async.bulk, shared ->
//            global, one RDM row segment each) and (b) feed the velocity CFAR (one thread per column half, window
//            cells summed directly from registers), 32 hit bits per thread, hits appended to the list.
//
// The 16 lanes of the wire format are interleaved at 4-byte granularity ([range][lane][I,Q]), so a one-lane CTA cannot
// read its input from the wire buffer efficiently.  Sixteen CTAs form a TEAM that works on the 16 lanes of one
// (CPI, tile) at a time.  Each member de-interleaves a sixteenth (4 of 64 PRTs) of the team's NEXT-BUT-ONE tile-group
// into a small ring in global memory (4 slots x 16 lanes x 68 KB per team, ~40 MB in total, written and re-read
// within microseconds, i.e. L2-resident), two rounds ahead of its use: every thread moves 8 pieces of [1 range][4
// lanes] per item, fetched with cp.async (16 bytes, global -> its private shared-memory staging slot, no register
// transit and therefore no exposed HBM latency), read back one pipeline point later and scattered as 4-byte stores
// to [lane][prt][range].  A member then fetches its lane's 64 x 256 samples with ONE bulk copy (cp.async.bulk, global
// -> shared, mbarrier complete_tx) while the previous item is still in phase 2.  Members publish "rounds produced"
// with st.release.gpu and poll their fifteen peers with ld.acquire.gpu (bounded); the grid is launched cooperatively
// so that all members are co-resident.
#include "common.cuh"
#include "radix.cuh"
#include "tw64.cuh"
#include "pc_core.cuh"
#include "kernels.h"
#include "../../include/radar_b200.h"
#include <algorithm>

namespace rb {

namespace op {
constexpr int kThreads = 512;
constexpr int kP = 64;                    // PRTs per CPI
constexpr int kLanes = 16;                // team size = interleaved lanes
constexpr int kNT = 256;                  // overlap-save tile
constexpr int kRowC = 256;                // complex slots per slab row (exchange layout 16 x 16, XOR-swizzled)
constexpr int kRowW = 2 * kRowC;          // 32-bit words per slab row
constexpr int kRawW = 272;                // words per staged raw row: 256 samples + 16 pad (consecutive rows 16 banks apart)
constexpr int kItemWords = kP * kRawW;    // one lane's input of an item in the ring
constexpr int kItemBytes = kItemWords * 4;            // 69 632
constexpr int kSlots = 4;                 // ring slots per team
constexpr int kSlabBytes = kP * kRowW * 4;            // 131 072
constexpr int kTabBytes = 64 * 8 + 64 * 4 + 2 * 256 * 8;                // T table [2][32] float2, window table [2][32], twiddles, spectrum
constexpr int kStageSlots = 3;            // cp.async staging: 16-byte pieces in flight per thread
constexpr int kStageBytes = kStageSlots * kThreads * 16;                // 24 576
constexpr int kSmemBytes = kSlabBytes + kItemBytes + kTabBytes + kStageBytes;    // 230 144 (+ the static mbarrier) of 232 448
constexpr int kPieces = 8;                // [1 range][4 lanes] pieces per thread per item (4 PRT x 256 range x 4 lane groups / 512)
constexpr int kMaxV = 192;
constexpr int kSpinLimit = 1 << 22;
}  // namespace op

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int4 ld_stream_v4(const int4* p) {
    int4 v;
    asm volatile("ld.global.cs.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// int16 pair -> (float I, float Q), exact, without I2F: 0x4B400000 | (s ^ 0x8000) is the float 12582912 + 32768 + s
__device__ __forceinline__ float2 unpack_iq(int w) {
    const unsigned t = (unsigned)w ^ 0x80008000u;
    const unsigned fi = __byte_perm(t, 0x4B400000u, 0x7610);
    const unsigned fq = __byte_perm(t, 0x4B400000u, 0x7632);
    return csub(make_float2(__uint_as_float(fi), __uint_as_float(fq)), make_float2(12615680.f, 12615680.f));
}

// 32-point forward DFT, natural order in and out, registers only: i = q + 8j, dft4 over j, twiddle w32^(q*k0), dft8 over q
__device__ __forceinline__ void fft32_fwd(float2 (&a)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        dft4<-1>(a[q], a[q + 8], a[q + 16], a[q + 24]);
#pragma unroll
        for (int k0 = 1; k0 < 4; ++k0) {
            const int m = 2 * ((q * k0) & 31);     // w32^(q k0) = w64^(2 q k0)
            if (m != 0) a[q + 8 * k0] = cmul(a[q + 8 * k0], make_float2(kCos64[m], -kSin64[m]));
        }
    }
    float2 f[32];
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) {
        float2 b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] = a[q + 8 * k0];
        dft8<-1>(b);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) f[k0 + 4 * k1] = b[k1];
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = f[k];
}

// velocity CA-CFAR for one column half (G = 0: rows 0..31, G = 1: rows 32..63) on dense magnitude rows in shared memory.
// Cropped axis y = row - 1 (n0 = 0: rows 1..63 are tested, row 0 is neither tested nor part of a window,
// CW/executeCFAR.m:23); windows y-12..y-8 and y+8..y+12 with the edge substitution of Function_CFAR1D_sub.m:30-39.
template <int G>
__device__ __forceinline__ unsigned cfar_half(const float* __restrict__ mg, int c, int meth_v, float tv_over_ref) {
    constexpr int REF = 5, GUARD = 7, NV = 63;
    constexpr int Y0 = G ? 31 : 0, Y1 = G ? 62 : 30;              // decided y (inclusive)
    constexpr int LO = (Y0 - GUARD - REF) < 0 ? 0 : (Y0 - GUARD - REF);
    constexpr int HI = (Y1 + GUARD + REF) > NV - 1 ? NV - 1 : (Y1 + GUARD + REF);
    float m[HI - LO + 1];
#pragma unroll
    for (int y = LO; y <= HI; ++y) {
        const int row = y + 1;
        m[y - LO] = mg[row * op::kRowW + 16 * ((row >> 4) & 1) + c];
    }
    unsigned hits = 0u;
#pragma unroll
    for (int y = Y0; y <= Y1; ++y) {
        const int l1 = y - GUARD - REF;
        const int r1 = y + GUARD + 1;
        const bool okL = l1 >= 0;
        const bool okR = r1 + REF - 1 <= NV - 1;
        float sl = 0.f, sr = 0.f;
#pragma unroll
        for (int j = 0; j < REF; ++j) {
            if (okL) sl += m[(okL ? l1 + j : LO) - LO];
            if (okR) sr += m[(okR ? r1 + j : LO) - LO];
        }
        const float a = okL ? sl : sr;
        const float b = okR ? sr : sl;
        const float mu = meth_v == 0 ? fmaxf(a, b) : fminf(a, b);
        if (m[y - LO] >= mu * tv_over_ref) hits |= 1u << (y + 1 - 32 * G);
    }
    return hits;
}

__device__ __forceinline__ void sts64(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }

__device__ __forceinline__ float op_fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- producer side: de-interleave PRTs 4*member .. 4*member+3 of tile-group (team + j*n_teams) into ring slot j & 3.
// Piece q (0..7) of thread t: T = t + 512*q -> lane group g = t & 3, range = ((t >> 2) & 127) + 128*(q & 1), prt_local =
// q >> 1; a warp's 32 pieces are 512 contiguous source bytes.  issue: cp.async 16 bytes into the thread's private staging
// slot; finish (one pipeline point later, after cp.async.wait_group): read it back and store lanes 4g..4g+3 to the ring.
struct ProdRound {          // per (thread, round) constants, set up once per item
    const int* src;         // wire word of (cpi, prt = 4*member, range = tile*V + ((t >> 2) & 127), lane 4g)
    int* dst;               // ring word of (slot, lane 4g, prt = 4*member, that range)
    uint32_t stage;         // shared address of the thread's staging slot 0
    int r;                  // tile*V + ((t >> 2) & 127)
    int R;
    bool on;
};
__device__ __forceinline__ ProdRound prod_round(const int* raw, int* ring_team, const int4* stage, int team, int member, int n_teams,
                                                int n_tiles, int R, int V, int t, int j, bool on) {
    ProdRound q;
    const int tg = team + j * n_teams;
    const int cpi = tg / n_tiles, tile = tg - cpi * n_tiles;
    const int g = t & 3, rng = (t >> 2) & 127;
    q.r = tile * V + rng;
    q.R = R;
    q.on = on;
    q.src = raw + (((size_t)(cpi * op::kP + member * 4) * R + q.r) * op::kLanes + 4 * g);
    q.dst = ring_team + ((size_t)(j & (op::kSlots - 1)) * op::kLanes + 4 * g) * op::kItemWords + member * 4 * op::kRawW + rng;
    q.stage = smem_u32(stage + t);
    return q;
}
template <int PIECE, int SLOT>
__device__ __forceinline__ void prod_issue(const ProdRound& q) {
    if (q.r + 128 * (PIECE & 1) < q.R) {
        const int* src = q.src + (size_t)(PIECE >> 1) * q.R * op::kLanes + (PIECE & 1) * 128 * op::kLanes;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(q.stage + SLOT * op::kThreads * 16), "l"(src) : "memory");
    }
}
template <int PIECE, int SLOT>
__device__ __forceinline__ void prod_finish(const ProdRound& q) {
    int4 a = make_int4(0, 0, 0, 0);                           // beyond the PRT: x = 0 (linear, not circular, correlation)
    if (q.r + 128 * (PIECE & 1) < q.R)
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(q.stage + SLOT * op::kThreads * 16) : "memory");
    int* dst = q.dst + (PIECE >> 1) * op::kRawW + (PIECE & 1) * 128;
    dst[0] = a.x;
    dst[op::kItemWords] = a.y;
    dst[2 * op::kItemWords] = a.z;
    dst[3 * op::kItemWords] = a.w;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// pipeline point N (0..3) of a round: finish the pieces issued at point N-1, issue the next ones (3, 3, 2 pieces)
template <int N>
__device__ __forceinline__ void prod_point(const ProdRound& q) {
    if (!q.on) return;
    if (N > 0) {
        cp_async_wait_all();
        prod_finish<3 * (N > 0 ? N - 1 : 0) + 0, 0>(q);
        prod_finish<3 * (N > 0 ? N - 1 : 0) + 1, 1>(q);
        if (3 * (N - 1) + 2 < op::kPieces) prod_finish<(3 * (N > 0 ? N - 1 : 0) + 2) & 7, 2>(q);
    }
    if (N < 3) {
        prod_issue<(3 * N + 0) & 7, 0>(q);
        prod_issue<(3 * N + 1) & 7, 1>(q);
        if (3 * N + 2 < op::kPieces) prod_issue<(3 * N + 2) & 7, 2>(q);
        cp_async_commit();
    }
}

// Phase 2, Doppler part (warps 0..11 own columns; the others only pass the barriers): a thread pair (lanes l, l ^ 16) owns column c; hh = 0 takes the even PRTs, hh = 1
// the odd ones pre-multiplied by (-1)^i, which rotates its 32-point spectrum by 16 bins so that both partners send their
// register index 16+i and keep index i in the radix-2 combine (no selects).
__device__ __noinline__ void doppler_phase2(float2* slab, const float* wtab, const float2* Ttab, int c, int hh, bool act, bool fft_warp,
                                               unsigned keep_bits, const ProdRound pr) {
    using namespace op;
    float* mg = reinterpret_cast<float*>(slab);
    if (!fft_warp) {                                // no column in this warp: only the barriers and the production point
        cta_sync();
        prod_point<3>(pr);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cta_sync();
        return;
    }
    // straight-line code from here on: the column stays in registers across the barrier
    float2 a[32];
    {
        const float2* col = slab + hh * kRowC + (act ? c : 0);
        const float* wt = wtab + 32 * hh;
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = cscale(col[2 * i * kRowC], wt[i]);   // x[2i + hh] * w  (MP/fun_Process_MTD.m:22)
    }
    cta_sync();                                     // every column is in registers: the slab rows become fp32 magnitude rows
    {
        fft32_fwd(a);
        const float2* tt = Ttab + 32 * hh;
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = cmul(a[i], tt[i]);
        float* mo = mg + hh * (16 * kRowW + 16) + c;    // rows 16*hh + i (and + 32), row offset 16*((row >> 4) & 1) = 16*hh words
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float2 rv;
            rv.x = __shfl_xor_sync(0xffffffffu, a[16 + i].x, 16);
            rv.y = __shfl_xor_sync(0xffffffffu, a[16 + i].y, 16);
            const float2 lo = cadd(a[i], rv);       // bin 16*hh + i        -> row 32 + 16*hh + i   (fftshift, :24)
            const float2 hi = csub(a[i], rv);       // bin 16*hh + i + 32   -> row 16*hh + i        (sign irrelevant under |.|)
            float mlo = op_fast_sqrt(lo.x * lo.x + lo.y * lo.y);
            float mhi = op_fast_sqrt(hi.x * hi.x + hi.y * hi.y);
            mlo = ((keep_bits >> (16 + i)) & 1u) ? mlo : 0.f;           // MP/fun_0v_pressing.m:4-6
            mhi = ((keep_bits >> i) & 1u) ? mhi : 0.f;
            if (act) {
                mo[(32 + i) * kRowW] = mlo;
                mo[i * kRowW] = mhi;
            }
        }
    }
    prod_point<3>(pr);                                                  // last pieces of the round: their cp.async had the FFT to land
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // magnitude rows (generic) before the bulk stores (async)
    cta_sync();
}

__global__ void __launch_bounds__(op::kThreads, 1) onepass_kernel(const __grid_constant__ OnePassParams p) {
    using namespace op;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* slab = reinterpret_cast<float2*>(smem_raw);                        // [64][256] complex; later [64] dense fp32 rows
    float* mg = reinterpret_cast<float*>(smem_raw);
    int* rawbuf = reinterpret_cast<int*>(smem_raw + kSlabBytes);               // [64][272] int16 pairs
    float2* Ttab = reinterpret_cast<float2*>(smem_raw + kSlabBytes + kItemBytes);   // [2][32]: 1 | w64^((k+16) mod 32)
    float* wtab = reinterpret_cast<float*>(Ttab + 64);                          // [2][32]: win[2i+h] * (h ? (-1)^i : 1)
    float2* tw_sm = reinterpret_cast<float2*>(wtab + 64);                       // [k][u]: w256^(u*k)
    float2* h_sm = tw_sm + 256;                                                 // [k][q]: spectrum bin q + 16*k
    int4* stage = reinterpret_cast<int4*>(smem_raw + kSlabBytes + kItemBytes + kTabBytes);   // [3][512] cp.async staging
    __shared__ __align__(8) uint64_t full_bar;

    const int t = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);     // warp-uniform by construction: branches on it do not diverge
    const int lane = t & 31;
    const int n1 = lane & 15, hh = lane >> 4;
    const int team = blockIdx.x / kLanes, member = blockIdx.x % kLanes;
    const int n_tg = p.n_cpi * p.n_tiles;
    if (team >= n_tg) return;                                                   // whole team leaves together
    const int n_rounds = (n_tg - team + p.n_teams - 1) / p.n_teams;
    const int R = p.R, V = p.V;

    if (t == 0) {
        mbar_init(&full_bar, 1);
        mbar_fence_init();
    }
    if (t < 64) {
        const int h = t >> 5, i = t & 31;
        wtab[t] = p.win[2 * i + h] * ((h && (i & 1)) ? -1.f : 1.f);
        const int m = (i + 16) & 31;
        Ttab[t] = h ? make_float2(kCos64[m], -kSin64[m]) : make_float2(1.f, 0.f);
    }
    if (t < 256) {
        tw_sm[t] = __ldg(p.tw + t);
        h_sm[t] = __ldg(p.hperm + (t & 15) * 16 + (t >> 4));               // transposed: h_sm[k*16 + q] = spectrum bin q + 16*k
    }
    // zero-velocity rows of this thread's Doppler outputs: bit i -> row 16*hh + i (bins +32), bit 16+i -> row 32 + 16*hh + i
    const unsigned keep_bits = (unsigned)((p.keep_mask >> (16 * hh)) & 0xffffull) | ((unsigned)((p.keep_mask >> (32 + 16 * hh)) & 0xffffull) << 16);

    int* const my_flag = p.flags + team * kLanes + member;
    const int* const team_flags = p.flags + team * kLanes;
    int* const ring_team = p.ring + (size_t)team * kSlots * kLanes * kItemWords;

    // ---- consumer side (warp 15): wait until every member has produced round j, then fetch this lane's input
    auto fetch = [&](int j) {
        int spins = 0;
        for (;;) {
            const int v = lane < kLanes ? ld_acquire_gpu(team_flags + lane) : 0x7fffffff;
            if (__all_sync(0xffffffffu, v >= j + 1)) break;
            if (++spins > kSpinLimit) {
                if (lane == 0) atomicExch(p.err_flag, 2);
                break;
            }
            __nanosleep(64);
        }
        if (lane == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");       // peers' generic-proxy stores -> this CTA's async-proxy read
            mbar_expect_tx(&full_bar, (uint32_t)kItemBytes);
            bulk_g2s(rawbuf, ring_team + ((size_t)(j & (kSlots - 1)) * kLanes + member) * kItemWords, (uint32_t)kItemBytes, &full_bar);
        }
    };

    // ---- prologue: produce rounds 0 and 1 (same pipeline points, back to back)
#pragma unroll 1
    for (int j = 0; j < 2 && j < n_rounds; ++j) {
        const ProdRound pr = prod_round(p.raw, ring_team, stage, team, member, p.n_teams, p.n_tiles, R, V, t, j, true);
        prod_point<0>(pr);
        prod_point<1>(pr);
        prod_point<2>(pr);
        prod_point<3>(pr);
    }
    __syncthreads();
    if (t == 0) {
        __threadfence();
        st_release_gpu(my_flag, 2);
    }
    if (warp == 15) fetch(0);

#pragma unroll 1
    for (int k = 0; k < n_rounds; ++k) {
        const int tg = team + k * p.n_teams;
        const int cpi = tg / p.n_tiles, tile = tg - cpi * p.n_tiles;
        const int r0 = tile * V;
        const int Vt = min(V, R - r0);
        const int slab_id = cpi * kLanes + member;

        // ================= phase 1: pulse compression of the 64 lines (two per warp at a time) =================
        // per-thread twiddles w256^(n1*k) of the pulse compression, register-resident for the 4 lines of this item (re-read
        // per item so they are not live across phase 2); the reference spectrum is used once per line and is read from
        // shared memory where it is needed (both half-warps read the same 128 bytes: one wavefront)
        float2 tw[15];
#pragma unroll
        for (int q = 1; q < 16; ++q) tw[q - 1] = tw_sm[q * 16 + n1];
        const float2* hq = h_sm + n1;
        const bool produce = k + 2 < n_rounds;
        // the per-round production constants are recomputed at every pipeline point instead of being kept live across the
        // register-tight transforms (jp is laundered through an empty asm so the compiler does not hoist them)
        auto round_now = [&]() {
            int jp = produce ? k + 2 : k;
            asm volatile("" : "+r"(jp));
            return prod_round(p.raw, ring_team, stage, team, member, p.n_teams, p.n_tiles, R, V, t, jp, produce);
        };
        mbar_wait(&full_bar, (uint32_t)(k & 1));
        prod_point<0>(round_now());
#pragma unroll 1
        for (int rho = 0; rho < 2; ++rho) {
            const int pl = 32 * rho + 2 * warp + hh;
            const int* rw = rawbuf + pl * kRawW + n1;
            // Exchange layout inside the line's own slab row (256 complex slots, 128-byte aligned): element (a, b) of the
            // 16 x 16 matrix lives at slot 16*a + (b ^ a).  Writers (thread b, instruction a) fill 16 consecutive slots,
            // readers (thread a, instruction b) hit 16 different bank pairs; with the row base 128-byte aligned the slot
            // address is (base | 8*thread) ^ 8*instr (+ 128*instr for writers): one LOP3 per access, immediates otherwise.
            const uint32_t rowS = smem_u32(slab + pl * kRowC);
            const uint32_t wA = rowS | (uint32_t)(n1 * 8);                       // writer: slot 16*q + (n1 ^ q)
            const uint32_t rB = (rowS + (uint32_t)(n1 * 128)) | (uint32_t)(n1 * 8);   // reader: slot 16*n1 + (m ^ n1)
            float2 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = unpack_iq(rw[16 * j]);          // x[n1 + 16 j]  (FrameDataRead_xzr.m:154-156)
            Dft<16, -1>::run(v);
#pragma unroll
            for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], tw[q - 1]);
#pragma unroll
            for (int q = 0; q < 16; ++q) sts64((wA ^ (uint32_t)(q * 8)) + q * 128, v[q]);      // element (k = q, n1)
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = lds64(rB ^ (uint32_t)(m * 8));                 // elements (k = n1, m)
            Dft<16, -1>::run(v);                                                // v[j] = X[n1 + 16 j]
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], hq[16 * j]);         // conj(FFT(taps)) * scale / 256 at bin n1 + 16 j
            Dft<16, +1>::run(v);
#pragma unroll
            for (int m = 1; m < 16; ++m) v[m] = cmulc(v[m], tw[m - 1]);
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) sts64((wA ^ (uint32_t)(m * 8)) + m * 128, v[m]);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = lds64(rB ^ (uint32_t)(q * 8));
            Dft<16, +1>::run(v);                                                // v[j] = y[n1 + 16 j]
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 12; ++j)
                if (n1 + 16 * j < V) sts64(rowS + (uint32_t)((n1 + 16 * j) * 8), v[j]);   // alias-free lags only (V <= 192)
            if (rho == 0) prod_point<1>(round_now());
            else prod_point<2>(round_now());
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // raw reads (generic) before the next bulk refill (async)
        __syncthreads();
        if (warp == 15 && k + 1 < n_rounds) fetch(k + 1);

        // ================= phase 2: Doppler FFT, |.|, 0-v mask, RDM rows out, velocity CFAR =================
        doppler_phase2(slab, wtab, Ttab, 16 * warp + n1, hh, 16 * warp + n1 < Vt, warp < 12 && 16 * warp < Vt, keep_bits, round_now());
        if (t < kP) {
            const int row = t;
            const float* src = mg + row * kRowW + 16 * ((row >> 4) & 1);
            float* dst = p.rdm + ((size_t)slab_id * kP + row) * R + r0;
            bulk_s2g(dst, src, (uint32_t)Vt * 4u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (t < 2 * kMaxV) {
            const int g = t >= kMaxV ? 1 : 0;
            const int cc = t - g * kMaxV;
            const bool cact = cc < Vt;
            unsigned hits = 0u;
            if (32 * ((t - g * kMaxV) >> 5) < Vt) {         // warp-uniform: any active column in this warp
                const int cl = cact ? cc : 0;
                hits = g ? cfar_half<1>(mg, cl, p.meth_v, p.tv_over_ref) : cfar_half<0>(mg, cl, p.meth_v, p.tv_over_ref);
                const int r = r0 + cc;
                int slo, shi;
                if (!cact || !cfar_seg_of(p.segs, r, R, &slo, &shi)) hits = 0u;
                if (cact) reinterpret_cast<unsigned*>(p.colmask)[((size_t)slab_id * R + r) * 2 + g] = hits;
                if (__any_sync(0xffffffffu, hits != 0u)) {
                    const int n = __popc(hits);
                    int incl = n;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int u = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += u;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    int base = 0;
                    if (lane == 31) base = atomicAdd(p.det_count, total);
                    base = __shfl_sync(0xffffffffu, base, 31);
                    int slot = base + incl - n;
                    unsigned hbits = hits;
                    while (hbits) {
                        const int row = __ffs((int)hbits) - 1 + 32 * g;
                        hbits &= hbits - 1;
                        if (slot < p.max_det) {
                            rb200_det d;
                            d.cpi = (uint32_t)(p.cpi0 + cpi);
                            d.r = (uint32_t)r;
                            d.v = (uint16_t)row;
                            d.lane = (uint8_t)member;
                            d.kind = RB200_DET_V;
                            d.amp = mg[row * kRowW + 16 * ((row >> 4) & 1) + cc];
                            reinterpret_cast<rb200_det*>(p.dets)[slot] = d;
                        }
                        ++slot;
                    }
                }
            }
        }
        if (t < kP) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the rows have left shared memory
        __syncthreads();                                    // slab free for the next item; this member's ring writes are done
        if (t == 0 && produce) {
            __threadfence();
            st_release_gpu(my_flag, k + 3);
        }
    }
}

// ---------------------------------------------------------------------------------------------
int onepass_tile_valid(int n_taps) {
    const int v = (op::kNT - n_taps + 1) & ~3;
    return std::min(v, op::kMaxV);
}

size_t onepass_ring_bytes(int n_teams) { return (size_t)n_teams * op::kSlots * op::kLanes * op::kItemBytes; }

int onepass_teams(int n_sms, int n_tile_groups) { return std::max(1, std::min(n_sms / op::kLanes, n_tile_groups)); }

cudaError_t launch_onepass(const OnePassParams& p, cudaStream_t st) {
    static size_t configured[64] = {};
    cudaError_t ce = ensure_dynamic_smem(onepass_kernel, (size_t)op::kSmemBytes, configured);
    if (ce != cudaSuccess) return ce;
    if (p.n_teams < 1 || p.n_cpi < 1 || p.n_tiles < 1 || p.V < 4 || p.V > op::kMaxV || (p.V & 3) || (p.R & 3)) return cudaErrorInvalidValue;
    void* args[] = {const_cast<OnePassParams*>(&p)};
    // cooperative launch: all CTAs of the grid are co-resident (the team hand-shake relies on it), and two such grids on
    // different streams are gang-scheduled one after the other instead of starving each other
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(onepass_kernel), dim3(p.n_teams * op::kLanes), dim3(op::kThreads), args,
                                       (size_t)op::kSmemBytes, st);
}

}  // namespace rb
