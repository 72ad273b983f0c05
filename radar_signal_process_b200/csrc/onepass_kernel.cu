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
// reference).  One CTA (384 threads = 12 warps, 168 registers, 1 per SM, ~226 KB of shared memory) owns an item:
//   phase 1  pulse compression of the 64 lines, WARP-PRIVATE and dynamically scheduled: a warp takes the next PAIR of
//            lines from a shared-memory counter and transforms them with 16 threads x 16 points per line (radix-16 x 16,
//            forward DIF -> reference spectrum -> inverse DIT); the two exchanges of a transform go through the line's
//            own row of the slab (in place, XOR-swizzled, __syncwarp only), so the warps drift through different phases
//            and the fp32 and shared-memory pipes overlap instead of alternating.  The 15 twiddles of a thread live in
//            registers in both operand forms a packed complex multiply needs ((x, y) and (-y, x): two FMUL2/FFMA2 per
//            product, no sign fix-ups); int16 -> fp32 uses the 2^23 magic-number construction (PRMT/LOP3 + one packed
//            FADD2; no I2F).
//   phase 2  six warps pull one Doppler column per thread out of the slab into registers (64 complex = 128 registers),
//            the CTA synchronises once, and from there on the column lives in registers only: window, 64-point FFT (8 x
//            radix-8, compile-time twiddles, 8 x radix-8), |.|, zero-velocity mask, RDM rows stored range-contiguous,
//            velocity CFAR on the register column, hit word + list append.  The OTHER six warps go straight on to phase 1
//            of the next item (the slab is free again), and the Doppler warps join them when their column is done: the
//            fp32-heavy Doppler work overlaps the shared-memory-heavy transforms of the next item.
//
// The 16 lanes of the wire format are interleaved at 4-byte granularity ([range][lane][I,Q]), so a one-lane CTA cannot
// read its input from the wire buffer efficiently.  Sixteen CTAs form a TEAM that works on the 16 lanes of one
// (CPI, tile) at a time.  Each member de-interleaves a sixteenth (4 of 64 PRTs) of the team's NEXT-BUT-ONE tile-group
// into a small ring in global memory (4 slots x 16 lanes x 68 KB per team, ~40 MB in total, written and re-read
// within microseconds, i.e. L2-resident): every thread moves up to 12 pieces of [1 range][4 lanes] per item, fetched
// with cp.async (16 bytes, global -> its private shared-memory staging slot, no register transit and therefore no
// exposed HBM latency), read back at the next pipeline point (after a line pair / around the Doppler column) and
// scattered as 4-byte stores to [lane][prt][range].  A member fetches its lane's input with two bulk copies (cp.async.bulk,
// global -> shared, mbarrier complete_tx; PRTs 0-31 and 32-63), each issued as soon as the sixteen line pairs of that
// half of the CURRENT item have been pulled into registers.  Members publish "rounds produced" with st.release.gpu and
// poll their fifteen peers with ld.acquire.gpu (bounded); the grid is launched cooperatively so that all members are
// co-resident.
#include "common.cuh"
#include "radix.cuh"
#include "tw64.cuh"
#include "pc_core.cuh"
#include "kernels.h"
#include "../../include/radar_b200.h"
#include <algorithm>

namespace rb {

namespace op {
constexpr int kThreads = 384;
constexpr int kDopplerWarps = 6;          // 6 x 32 columns >= V
constexpr int kP = 64;                    // PRTs per CPI
constexpr int kPairs = kP / 2;            // line pairs per item
constexpr int kLanes = 16;                // team size = interleaved lanes
constexpr int kNT = 256;                  // overlap-save tile
constexpr int kRowC = 256;                // complex slots per slab row (exchange layout 16 x 16, XOR-swizzled)
constexpr int kRawW = 272;                // words per staged raw row: 256 samples + 16 pad (consecutive rows 16 banks apart)
constexpr int kItemWords = kP * kRawW;    // one lane's input of an item in the ring
constexpr int kHalfBytes = kItemWords * 2;            // 34 816: PRTs 0-31 or 32-63
constexpr int kSlots = 4;                 // ring slots per team
constexpr int kSlabBytes = kP * kRowC * 8;            // 131 072
constexpr int kRawBytes = kItemWords * 4;             // 69 632
constexpr int kTabBytes = 256 * 8 + 256 * 16;         // twiddles (x, y); spectrum in both operand forms (x, y, -y, x)
constexpr int kStageSlots = 4;            // cp.async staging: 16-byte pieces in flight per thread
constexpr int kStageBytes = kStageSlots * kThreads * 16;                // 24 576
constexpr int kSmemBytes = kSlabBytes + kRawBytes + kTabBytes + kStageBytes;     // 231 424 (+ 48 static) of 232 448
constexpr int kMaxV = 32 * kDopplerWarps; // 192
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

// int16 pair -> (float I, float Q), exact, without I2F: 0x4B400000 | (s ^ 0x8000) is the float 12582912 + 32768 + s
__device__ __forceinline__ float2 unpack_iq(int w) {
    const unsigned t = (unsigned)w ^ 0x80008000u;
    const unsigned fi = __byte_perm(t, 0x4B400000u, 0x7610);
    const unsigned fq = __byte_perm(t, 0x4B400000u, 0x7632);
    return csub(make_float2(__uint_as_float(fi), __uint_as_float(fq)), make_float2(12615680.f, 12615680.f));
}

// complex products with the constant operand b held in BOTH forms P = (b.x, b.y), M = (-b.y, b.x): two packed
// instructions each, the half swaps are free operand modifiers (.LO_HI)
#if defined(RB_PACKED_F32)
__device__ __forceinline__ float2 cmul_pm(float2 a, float2 P, float2 M) {            // a * b
    return rb_up(rb_fma2(rb_pk(a.x, a.x), rb_pk(P.x, P.y), rb_mul2(rb_pk(a.y, a.y), rb_pk(M.x, M.y))));
}
__device__ __forceinline__ float2 cmulc_pm(float2 a, float2 P, float2 M) {           // a * conj(b) = a.x (b.x, -b.y) + a.y (b.y, b.x)
    return rb_up(rb_fma2(rb_pk(a.x, a.x), rb_pk(M.y, M.x), rb_mul2(rb_pk(a.y, a.y), rb_pk(P.y, P.x))));
}
#else       // host compilation pass (and RB_NO_PACKED_F32 builds): the same products in scalar form
__device__ __forceinline__ float2 cmul_pm(float2 a, float2 P, float2 M) { (void)M; return cmul(a, P); }
__device__ __forceinline__ float2 cmulc_pm(float2 a, float2 P, float2 M) { (void)M; return cmulc(a, P); }
#endif

// ---- producer side: de-interleave PRTs 4*member .. 4*member+3 of tile-group (team + j*n_teams) into ring slot j & 3.
// Thread t owns lane group g = t & 3 and ranges (t >> 2) + 96*sub, sub = 0..2 (< 256), of each of the four PRTs: piece
// p = 3*prt_local + sub, 12 pieces in three groups of four; a warp's 32 pieces of one (prt, sub) are 512 contiguous source
// bytes.  issue: cp.async 16 bytes into the thread's private staging slot; finish (a pipeline point later, after
// cp.async.wait_group): read it back and store lanes 4g..4g+3 to the ring.
struct ProdRound {          // per (thread, round) constants
    const int* src;         // wire word of (cpi, prt = 4*member, range = tile*V + (t >> 2), lane 4g)
    int* dst;               // ring word of (slot, lane 4g, prt = 4*member, that range)
    uint32_t stage;         // shared address of the thread's staging slot 0
    int r;                  // tile*V + (t >> 2)
    int rb;                 // t >> 2
    int R;
};
__device__ __forceinline__ ProdRound prod_round(const int* raw, int* ring_team, uint32_t stage_t, int team, int member, int n_teams,
                                                int n_tiles, int R, int V, int t, int j) {
    ProdRound q;
    const int tg = team + j * n_teams;
    const int cpi = tg / n_tiles, tile = tg - cpi * n_tiles;
    const int g = t & 3;
    q.rb = t >> 2;
    q.r = tile * V + q.rb;
    q.R = R;
    q.src = raw + (((size_t)(cpi * op::kP + member * 4) * R + q.r) * op::kLanes + 4 * g);
    q.dst = ring_team + ((size_t)(j & (op::kSlots - 1)) * op::kLanes + 4 * g) * op::kItemWords + member * 4 * op::kRawW + q.rb;
    q.stage = stage_t;
    return q;
}
template <int PIECE, int SLOT>
__device__ __forceinline__ void prod_issue(const ProdRound& q) {
    constexpr int PL = PIECE / 3, SUB = PIECE % 3;
    if (q.rb + 96 * SUB < 256 && q.r + 96 * SUB < q.R) {
        const int* src = q.src + (size_t)PL * q.R * op::kLanes + SUB * 96 * op::kLanes;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(q.stage + SLOT * op::kThreads * 16), "l"(src) : "memory");
    }
}
template <int PIECE, int SLOT>
__device__ __forceinline__ void prod_finish(const ProdRound& q) {
    constexpr int PL = PIECE / 3, SUB = PIECE % 3;
    if (q.rb + 96 * SUB >= 256) return;
    int4 a = make_int4(0, 0, 0, 0);                           // beyond the PRT: x = 0 (linear, not circular, correlation)
    if (q.r + 96 * SUB < q.R)
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(q.stage + SLOT * op::kThreads * 16) : "memory");
    int* dst = q.dst + PL * op::kRawW + SUB * 96;
    dst[0] = a.x;
    dst[op::kItemWords] = a.y;
    dst[2 * op::kItemWords] = a.z;
    dst[3 * op::kItemWords] = a.w;
}
template <int G>
__device__ __forceinline__ void prod_issue_group(const ProdRound& q) {
    prod_issue<4 * G + 0, 0>(q);
    prod_issue<4 * G + 1, 1>(q);
    prod_issue<4 * G + 2, 2>(q);
    prod_issue<4 * G + 3, 3>(q);
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int G>
__device__ __forceinline__ void prod_finish_group(const ProdRound& q) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    prod_finish<4 * G + 0, 0>(q);
    prod_finish<4 * G + 1, 1>(q);
    prod_finish<4 * G + 2, 2>(q);
    prod_finish<4 * G + 3, 3>(q);
}
// One pipeline point of a round.  st: 0 = nothing issued yet, 1..3 = group st-1 in flight, 4 = round complete.
__device__ __forceinline__ void prod_step(const ProdRound& q, int& st) {
    switch (st) {
        case 0: prod_issue_group<0>(q); st = 1; break;
        case 1: prod_finish_group<0>(q); prod_issue_group<1>(q); st = 2; break;
        case 2: prod_finish_group<1>(q); prod_issue_group<2>(q); st = 3; break;
        case 3: prod_finish_group<2>(q); st = 4; break;
        default: break;
    }
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

__global__ void __launch_bounds__(op::kThreads, 1) onepass_kernel(const __grid_constant__ OnePassParams p) {
    using namespace op;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* slab = reinterpret_cast<float2*>(smem_raw);                        // [64][256] complex, rows XOR-swizzled during a transform
    int* rawbuf = reinterpret_cast<int*>(smem_raw + kSlabBytes);               // [64][272] int16 pairs (two halves of 32 rows)
    float2* tw_sm = reinterpret_cast<float2*>(smem_raw + kSlabBytes + kRawBytes);    // [k][u]: w256^(u*k)
    float4* h_sm = reinterpret_cast<float4*>(tw_sm + 256);                      // [k][q]: spectrum bin q + 16*k as (x, y, -y, x)
    int4* stage = reinterpret_cast<int4*>(smem_raw + kSlabBytes + kRawBytes + kTabBytes);   // [4][384] cp.async staging
    __shared__ __align__(8) uint64_t full_bar[2];      // raw half h of the current item has landed
    __shared__ __align__(8) uint64_t empty_bar[2];     // the 16 line pairs of raw half h have been pulled into registers
    __shared__ int pair_ctr[2];                        // next line pair of item k (slot k & 1)
    __shared__ int seen_rounds;                        // rounds every team member is known to have produced (monotonic cache)

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
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_init(&empty_bar[0], kPairs / 2);
        mbar_init(&empty_bar[1], kPairs / 2);
        mbar_fence_init();
        pair_ctr[0] = pair_ctr[1] = 0;
        seen_rounds = 0;
    }
    if (t < 256) {
        tw_sm[t] = __ldg(p.tw + t);
        const float2 h = __ldg(p.hperm + (t & 15) * 16 + (t >> 4));               // transposed: h_sm[k*16 + q] = bin q + 16*k
        h_sm[t] = make_float4(h.x, h.y, -h.y, h.x);
    }

    int* const my_flag = p.flags + team * kLanes + member;
    const int* const team_flags = p.flags + team * kLanes;
    int* const ring_team = p.ring + (size_t)team * kSlots * kLanes * kItemWords;
    const uint32_t stage_t = smem_u32(stage + t);

    // ---- consumer side (any one warp): wait until every member has produced round j, then fetch half `half` of this lane's input
    auto fetch_half = [&](int j, int half) {
        int spins = 0;
        // the members run about two rounds ahead, so one poll usually covers the next fetches as well: remember the minimum
        while (*reinterpret_cast<volatile int*>(&seen_rounds) < j + 1) {
            const int v = lane < kLanes ? ld_acquire_gpu(team_flags + lane) : 0x7fffffff;
            const int vmin = __reduce_min_sync(0xffffffffu, v);
            if (vmin >= j + 1) {
                if (lane == 0) atomicMax(&seen_rounds, vmin);
                break;
            }
            if (++spins > kSpinLimit) {
                if (lane == 0) atomicExch(p.err_flag, 2);
                break;
            }
            __nanosleep(64);
        }
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");       // peers' generic-proxy stores -> this CTA's async-proxy read
            mbar_expect_tx(&full_bar[half], (uint32_t)kHalfBytes);
            bulk_g2s(rawbuf + half * (kItemWords / 2),
                     ring_team + ((size_t)(j & (kSlots - 1)) * kLanes + member) * kItemWords + half * (kItemWords / 2), (uint32_t)kHalfBytes,
                     &full_bar[half]);
        }
    };
    auto round_of = [&](int j) { return prod_round(p.raw, ring_team, stage_t, team, member, p.n_teams, p.n_tiles, R, V, t, j); };

    // ---- prologue: produce rounds 0 and 1 back to back, publish, fetch item 0
#pragma unroll 1
    for (int j = 0; j < 2 && j < n_rounds; ++j) {
        const ProdRound pr = round_of(j);
        int st = 0;
#pragma unroll 1
        while (st < 4) prod_step(pr, st);
    }
    __syncthreads();
    if (t == 0) {
        __threadfence();
        st_release_gpu(my_flag, 2);
    }
    if (warp == 0) {
        fetch_half(0, 0);
        fetch_half(0, 1);
    }
    // production window W(k) = (barrier B of item k-1, barrier A of item k) moves round k + 2
    int pst = n_rounds > 2 ? 0 : 4;
    int pround = 2;
    ProdRound pr = round_of(n_rounds > 2 ? 2 : 0);

    const bool tracing = p.trace != nullptr && blockIdx.x == 0;
    auto stamp = [&](int k, int ev, long long val) {
        if (tracing && lane == 0 && k < 32) p.trace[((size_t)k * 12 + warp) * 16 + ev] = (unsigned long long)val;
    };
#pragma unroll 1
    for (int k = 0; k < n_rounds; ++k) {
        const int tg = team + k * p.n_teams;
        const int cpi = tg / p.n_tiles, tile = tg - cpi * p.n_tiles;
        const int r0 = tile * V;
        const int Vt = min(V, R - r0);
        const int slab_id = cpi * kLanes + member;

        // ================= phase 1: pulse compression, one line pair at a time per warp =================
        {
            // this thread's twiddles w256^(n1*q) in both operand forms, register-resident while the warp takes pairs
            float2 twP[15], twM[15];
#pragma unroll
            for (int q = 1; q < 16; ++q) {
                twP[q - 1] = tw_sm[q * 16 + n1];
                twM[q - 1] = make_float2(-twP[q - 1].y, twP[q - 1].x);
                // opaque to the optimiser: otherwise ptxas rematerialises (MOV + FADD) this operand in front of every product
                asm volatile("" : "+f"(twM[q - 1].x), "+f"(twM[q - 1].y));
            }
            const float4* hq = h_sm + n1;
            stamp(k, 0, clock64());
            int n_pairs = 0;
            long long t_wait = 0, t_fetch = 0;
            // the index of the NEXT pair is requested while the current one is being transformed (the shared-memory atomic and
            // the broadcast are off the critical path); a raw half is waited for once per warp and item
            int pair = 0;
            if (lane == 0) pair = atomicAdd(&pair_ctr[k & 1], 1);
            pair = __shfl_sync(0xffffffffu, pair, 0);
            unsigned ready = 0u;
#pragma unroll 1
            while (pair < kPairs) {
                int next = 0;
                if (lane == 0) next = atomicAdd(&pair_ctr[k & 1], 1);
                const int half = pair >> 4;
                const long long tw0 = tracing ? clock64() : 0;
                if (!((ready >> half) & 1u)) {
                    mbar_wait(&full_bar[half], (uint32_t)(k & 1));
                    ready |= 1u << half;
                }
                if (tracing) {
                    t_wait += clock64() - tw0;
                    if (n_pairs == 0) stamp(k, 1, clock64());
                    ++n_pairs;
                }
                const int pl = 2 * pair + hh;
                const int* rw = rawbuf + pl * kRawW + n1;
                float2 v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = unpack_iq(rw[16 * j]);      // x[n1 + 16 j]  (FrameDataRead_xzr.m:154-156)
                // this pair's raw rows are in registers: when the 16 pairs of a half are, the half is refilled for item k + 1
                // (generic-proxy reads -> fence -> mbarrier arrive [release]; the warp holding the half's last pair waits for the
                // sixteen arrivals [acquire] and issues the bulk copy [async proxy])
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[half]);
                if ((pair & 15) == 15) {
                    mbar_wait(&empty_bar[half], (uint32_t)(k & 1));
                    if (k + 1 < n_rounds) {
                        const long long tf0 = tracing ? clock64() : 0;
                        fetch_half(k + 1, half);
                        if (tracing) { t_fetch += clock64() - tf0; stamp(k, 10 + half, clock64()); }
                    }
                }

                // Exchange layout inside the line's own slab row (256 complex slots, 128-byte aligned): element (a, b) of the
                // 16 x 16 matrix lives at slot 16*a + (b ^ a).  Writers (thread b, instruction a) fill 16 consecutive slots,
                // readers (thread a, instruction b) hit 16 different bank pairs; with the row base 128-byte aligned the slot
                // address is (base | 8*thread) ^ 8*instr (+ 128*instr for writers): one LOP3 per access, immediates otherwise.
                const uint32_t rowS = smem_u32(slab + pl * kRowC);
                const uint32_t wA = rowS | (uint32_t)(n1 * 8);                       // writer: slot 16*q + (n1 ^ q)
                const uint32_t rB = (rowS + (uint32_t)(n1 * 128)) | (uint32_t)(n1 * 8);   // reader: slot 16*n1 + (m ^ n1)
                if (p.dbg & 2) {                                                // timing experiment: no transforms
                    if (pst < 4 && !(p.dbg & 4)) prod_step(pr, pst);
                    pair = __shfl_sync(0xffffffffu, next, 0);
                    continue;
                }
                Dft<16, -1>::run(v);
#pragma unroll
                for (int q = 1; q < 16; ++q) v[q] = cmul_pm(v[q], twP[q - 1], twM[q - 1]);
#pragma unroll
                for (int q = 0; q < 16; ++q) sts64((wA ^ (uint32_t)(q * 8)) + q * 128, v[q]);      // element (k = q, n1)
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; ++m) v[m] = lds64(rB ^ (uint32_t)(m * 8));                 // elements (k = n1, m)
                Dft<16, -1>::run(v);                                            // v[j] = X[n1 + 16 j]
#pragma unroll
                for (int j = 0; j < 16; ++j) {                                  // conj(FFT(taps)) * scale / 256 at bin n1 + 16 j
                    const float4 h = hq[16 * j];
                    v[j] = cmul_pm(v[j], make_float2(h.x, h.y), make_float2(h.z, h.w));
                }
                Dft<16, +1>::run(v);
#pragma unroll
                for (int m = 1; m < 16; ++m) v[m] = cmulc_pm(v[m], twP[m - 1], twM[m - 1]);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; ++m) sts64((wA ^ (uint32_t)(m * 8)) + m * 128, v[m]);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = lds64(rB ^ (uint32_t)(q * 8));
                Dft<16, +1>::run(v);                                            // v[j] = y[n1 + 16 j]
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 12; ++j)
                    if (n1 + 16 * j < V) sts64(rowS + (uint32_t)((n1 + 16 * j) * 8), v[j]);   // alias-free lags only (V <= 192)
                // pipeline point of the de-interleave
                if (pst < 4 && !(p.dbg & 4)) prod_step(pr, pst);
                pair = __shfl_sync(0xffffffffu, next, 0);
            }
            stamp(k, 8, n_pairs);
            stamp(k, 9, t_wait);
            stamp(k, 12, t_fetch);
        }
        if (p.dbg & 4) pst = 4;
        stamp(k, 2, clock64());
        // flush this thread's share of the round (normally one group is left in flight)
#pragma unroll 1
        while (pst < 4) prod_step(pr, pst);
        stamp(k, 3, clock64());
        cta_sync();                                         // barrier A: the slab holds the 64 compressed lines; round k + 2 is in the ring
        if (t == 0) {
            if (k + 2 < n_rounds) {
                __threadfence();
                st_release_gpu(my_flag, k + 3);
            }
            pair_ctr[(k + 1) & 1] = 0;
        }
        stamp(k, 4, clock64());
        pround = k + 3;
        pst = pround < n_rounds ? 0 : 4;
        if (pst < 4) pr = round_of(pround);

        // ================= phase 2: Doppler columns in registers; the other warps go on to the next item =================
        if (warp < kDopplerWarps) {
            const int c = 32 * warp + lane;
            const bool ok = c < Vt;
            float2 v[64];
            {
                const float2* col = slab + (ok ? c : Vt - 1);
#pragma unroll
                for (int prt = 0; prt < 64; ++prt) v[prt] = cscale(col[prt * kRowC], p.win[prt]);   // MP/fun_Process_MTD.m:22
            }
            stamp(k, 5, clock64());
            cta_sync();                                     // barrier B: every column is in registers, the slab is free again
            stamp(k, 6, clock64());
            if (pst < 4) prod_step(pr, pst);
            if (!(p.dbg & 1)) doppler_column(v, p, slab_id, cpi, member, r0 + (ok ? c : Vt - 1), ok, lane);
            stamp(k, 7, clock64());
            if (pst < 4) prod_step(pr, pst);
        } else {
            stamp(k, 5, clock64());
            cta_sync();                                     // barrier B
            stamp(k, 6, clock64());
            if (pst < 4) prod_step(pr, pst);
        }
    }
}

// ---------------------------------------------------------------------------------------------
int onepass_tile_valid(int n_taps) {
    const int v = (op::kNT - n_taps + 1) & ~3;
    return std::min(v, op::kMaxV);
}

size_t onepass_ring_bytes(int n_teams) { return (size_t)n_teams * op::kSlots * op::kLanes * op::kItemWords * 4; }

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
