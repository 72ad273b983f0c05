// api.cu -- context, plans and the extern "C" entry points of libradar_b200.so (include/radar_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/radar_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace rb;
typedef std::complex<double> cd;

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// iterative radix-2 FFT in double (host, plan-time only; n is a power of two)
void host_fft_pow2(std::vector<cd>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double)len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const cd w = std::polar(1.0, ang * (double)k);
                const cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

long mround(double x) { return (long)(x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5)); }

// kaiser(n, beta) of the Signal Processing Toolbox (MP/fun_Process_MTD.m:13-14)
std::vector<double> kaiser_window(int n, double beta) {
    std::vector<double> w(n, 1.0);
    if (n <= 1) return w;
    const double denom = std::cyl_bessel_i(0.0, std::fabs(beta));
    const double alpha = (n - 1) / 2.0;
    for (int k = 0; k < n; ++k) {
        const double r = (k - alpha) / alpha;
        w[k] = std::cyl_bessel_i(0.0, std::fabs(beta) * std::sqrt(std::max(0.0, 1.0 - r * r))) / denom;
    }
    return w;
}

// round(mean(grpdelay(b))) for a real FIR (MTD/fun_lss_pulse_compression.m:47), 512 points on [0,pi)
int grpdelay_mean_round(const std::vector<double>& b) {
    double acc = 0, babs = 0;
    for (double x : b) babs += std::fabs(x);
    for (int i = 0; i < 512; ++i) {
        const double w = M_PI * i / 512.0;
        cd num = 0, den = 0;
        for (size_t k = 0; k < b.size(); ++k) {
            const cd e = std::polar(1.0, -w * (double)k);
            num += (double)k * b[k] * e;
            den += b[k] * e;
        }
        if (std::abs(den) < 10 * 2.220446049250313e-16 * std::max(1.0, babs)) continue;   // singular -> 0
        acc += (num / den).real();
    }
    return (int)mround(acc / 512.0);
}

struct SegPlan {
    PcSegDev d;
    int nt;     // FFT tile (0 = time-domain fallback)
};

struct ClassPlan {   // all segments sharing one FFT tile size
    int nt = 0;
    int n_tiles = 0;
    DevBuf tiles, tw;
};

struct Plan {
    std::vector<SegPlan> segs;
    std::vector<ClassPlan> classes;
    DevBuf hperm, taps;
    int h_entries = 0;             // float2 entries in hperm
    std::vector<std::pair<int, int>> out_ranges;   // sorted [start, end) written by segments
    int max_in_end = 0, max_out_end = 0;
    bool valid = false;
    void release() {
        for (auto& c : classes) { c.tiles.release(); c.tw.release(); }
        classes.clear();
        hperm.release();
        taps.release();
        segs.clear();
        out_ranges.clear();
        max_in_end = max_out_end = 0;
        valid = false;
    }
};

struct MtdPlan {
    int P = 0;
    double beta = 0;
    std::vector<float> h_window;
    DevBuf window, tw;
    DevBuf tc_mat;              // RB200_MTD_TC: split bf16 DFT matrix of mtd64_tc_kernel for the zero-velocity rows below
    int tc_zlo = -2, tc_zhi = -2;
    int n_stages = 0;
    int radix[16];
};

}  // namespace

// Experiment switches, read from the environment ONCE in rb200_create -- never on a call path.
struct EnvSwitches {
    int pc_nt = 0;              // RB200_PC_NT
    int chunk = 0;              // RB200_CHUNK
    int slots = 0;              // RB200_SLOTS (0 = default)
    bool no_tma = false, no_tma_mtd = false, no_fused = false, no_fused_v = false, mega = false, no_cfar_tile = false;
    bool coexist = false;       // RB200_COEXIST=1: 12-warp pcw_kernel + one mtd64_tma CTA per SM, so that K1 of chunk i+1 and K2 of chunk i share the SMs
    bool mtd_tc = false;        // RB200_MTD_TC=1 (with RB200_NO_FUSED=1): P = 64 Doppler transform on the tensor cores (experiment, mtd64_tc_kernel.cu)
    int split = 0;              // RB200_SPLIT=n: pcw_kernel on n SMs and mtd64_tma on the others at the same time, K2 consuming each CPI as soon
                                // as K1 has finished it (the intermediate is then read from L2); experiment, device-resident batches only
    bool no_pcw = false;        // RB200_NO_PCW=1: the CTA-wide pc_fft_tma_kernel instead of the warp-private pcw_kernel
    bool onepass = false;       // RB200_ONEPASS=1: the single-pass kernel (PC intermediate in shared memory, onepass_kernel.cu)
    int op_dbg = 0;             // RB200_OP_DBG: timing experiments of the single-pass kernel (results are wrong)
    void read() {
        auto flag = [](const char* n) { const char* v = getenv(n); return v != nullptr && v[0] != 0 && !(v[0] == '0' && v[1] == 0); };
        auto num = [](const char* n) { const char* v = getenv(n); return v ? atoi(v) : 0; };
        pc_nt = num("RB200_PC_NT");
        chunk = num("RB200_CHUNK");
        slots = num("RB200_SLOTS");
        no_tma = flag("RB200_NO_TMA");
        no_pcw = flag("RB200_NO_PCW");
        coexist = flag("RB200_COEXIST");
        mtd_tc = flag("RB200_MTD_TC");
        split = num("RB200_SPLIT");
        no_tma_mtd = flag("RB200_NO_TMA_MTD");
        no_fused = flag("RB200_NO_FUSED");
        no_fused_v = flag("RB200_NO_FUSED_V");
        mega = flag("RB200_MEGA");
        no_cfar_tile = flag("RB200_NO_CFAR_TILE");
        onepass = flag("RB200_ONEPASS");
        op_dbg = num("RB200_OP_DBG");
    }
};

struct rb200_ctx {
    int device = 0;
    EnvSwitches env;
    cudaStream_t stream = nullptr;
    rb200_config cfg;
    std::string err;
    Plan plan;
    CfarSegs cfar_segs = {};       // rb200_set_cfar_segments (fun_CFARflag); n = 0: whole range axis
    Plan dmx_plan;                 // private plan of rb200_dmx_process_z (never touches the caller's waveform)
    unsigned long long dmx_key = 0;
    uint64_t plan_tag = 0;         // rb200_set_plan_tag / rb200_get_plan_tag; cleared by rb200_set_waveform
    Plan pcz_plan;                 // cached plan of rb200_pulse_compression_z, keyed on (L, M, hash of the taps)
    unsigned long long pcz_key = 0;
    std::map<std::pair<int, long long>, MtdPlan*> mtd_plans;
    DevBuf gain;
    int gain_n = 0;
    // chain buffers
    DevBuf raw, pc, rdm, dets_v, dets_2d, counters, vmask, errflag, colmask;
    DevBuf dbf_w;                  // DBF weights float2 [beam][channel]; dbf_beams = 0 when off
    int dbf_beams = 0;
    DevBuf split_ctr;              // RB200_SPLIT: per-chunk progress counters and start flags
    DevBuf ring, megactr;          // fused persistent chain: L2-resident PC ring, work / completion counters
    bool last_was_mega = false;
    bool last_was_onepass = false;
    bool keep_pc = false;          // rb200_set_debug_keep_pc: run the chain with the pulse-compressed intermediate in HBM
    int coop_launch = 0;           // cudaDevAttrCooperativeLaunch
    int n_sms = 148;
    // persistent-grid sizing (CTAs per SM): 3/3 fills the SM with one kernel at a time; 2/1 lets the compute-bound PC
    // kernel of chunk i+1 and the HBM-bound MTD kernel of chunk i be co-resident (registers: 2*20.5K + 19.6K <= 64K)
    int pc_ctas_per_sm = 4, mtd_ctas_per_sm = 3;
    // chunk pipelining of the fused path: chunk i runs on slot i % n_slots (own stream + scratch), so the
    // tail of one chunk's kernels overlaps the head of the next while the PC intermediate stays L2-sized
    struct Slot {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf pc, colmask, vlist, count, raw, rdm, beams;
        DevBuf op_planar;              // single-pass kernel: de-interleaved lane planes of this slot's chunk
        size_t op_planar_cap = 0;      // geometry the pad columns of op_planar were initialised for
        int op_R = 0, op_V = 0;
    };
    static const int kMaxSlots = 4;
    Slot slots[kMaxSlots];
    cudaEvent_t fork_ev = nullptr;
    const float2* last_pc = nullptr;
    cudaStream_t last_stream = nullptr;   // stream of the last chain_enqueue: rb200_chain_fetch orders itself after it
    // MATLAB-layout scratch
    DevBuf s_in_re, s_in_im, s_a, s_b, s_c, s_out_re, s_out_im, s_u8a, s_u8b, s_idx;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_timing = false;
    int launches = 0;
    int last_chunk_cpis = 0;
    // pinned staging for host detections
    rb200_det* h_dets = nullptr;
    int* h_counts = nullptr;
    // optional per-stage timing (bench.py roofline): 4 events per chunk, consumed by rb200_get_stage_ms
    bool stage_timing = false;
    std::vector<cudaEvent_t> stage_events;
    size_t stage_used = 0;
    std::vector<int> stage_cpis;       // CPIs per timed chunk
    int stage_k1_launches = 0;
};

static cudaEvent_t stage_event(rb200_ctx* c, cudaStream_t st) {
    if (!c->stage_timing || c->stage_used >= 65536) return nullptr;
    if (c->stage_used == c->stage_events.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        c->stage_events.push_back(e);
    }
    cudaEvent_t e = c->stage_events[c->stage_used++];
    cudaEventRecord(e, st);
    return e;
}

static std::string g_create_error;

#define CK(ctx, call)                                                                                 \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            char buf__[512];                                                                          \
            snprintf(buf__, sizeof buf__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            (ctx)->err = buf__;                                                                       \
            return RB200_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

static int fail(rb200_ctx* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}

// ---------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------
static int choose_nt(int L, int forced_nt) {
    if (forced_nt) {
        const int nt = forced_nt;
        if ((nt == 256 || nt == 512 || nt == 4096) && nt - L + 1 >= nt / 8) return nt;
    }
    // estimated butterfly flops per transformed point (forward + inverse + spectrum multiply)
    const int cand[3] = {256, 512, 4096};
    const double cost[3] = {59.0, 75.0, 92.0};
    int best = 0;
    double bestc = 1e300;
    for (int i = 0; i < 3; ++i) {
        const int V = cand[i] - L + 1;
        if (V < cand[i] / 8) continue;
        const double c = cost[i] * cand[i] / V;
        if (c < bestc) { bestc = c; best = cand[i]; }
    }
    return best;
}

static int build_plan(rb200_ctx* ctx, Plan& plan, const rb200_segment* segs, int nseg) {
    plan.release();
    if (nseg < 1 || nseg > kMaxSegs || !segs) return fail(ctx, RB200_ERR_ARG, "set_waveform: need 1..8 segments");
    std::vector<std::vector<cd>> corr_taps(nseg);   // t[k] of the correlation form
    std::vector<float2> taps_all;
    size_t h_total = 0;
    for (int i = 0; i < nseg; ++i) {
        const rb200_segment& s = segs[i];
        if (s.n_taps < 1 || !s.taps_re || s.in_len < 0 || s.out_len < 0 || s.in_start < 0 || s.out_start < 0)
            return fail(ctx, RB200_ERR_ARG, "set_waveform: bad segment");
        const int L = s.n_taps;
        std::vector<cd> raw(L);
        for (int k = 0; k < L; ++k) raw[k] = cd(s.taps_re[k], s.taps_im ? s.taps_im[k] : 0.0);
        SegPlan sp;
        memset(&sp, 0, sizeof sp);
        sp.d.in_start = s.in_start;
        sp.d.in_len = s.in_len;
        sp.d.out_start = s.out_start;
        sp.d.out_len = s.out_len;
        sp.d.n_taps = L;
        if (s.kind == RB200_SEG_MF) {
            if (s.align != RB200_ALIGN_LEADING_EDGE) return fail(ctx, RB200_ERR_ARG, "set_waveform: MF segments use LEADING_EDGE");
            corr_taps[i] = raw;                       // y[n] = sum x[n+k] conj(s0[k])
            sp.d.pre = 0;
            sp.d.rot = 0;
        } else if (s.kind == RB200_SEG_FIR) {
            corr_taps[i].resize(L);                   // y[n] = sum_j x[n-(L-1)+j] b[L-1-j]
            for (int j = 0; j < L; ++j) corr_taps[i][j] = std::conj(raw[L - 1 - j]);
            sp.d.pre = L - 1;
            sp.d.rot = 0;
            if (s.align == RB200_ALIGN_GRPDELAY) {
                std::vector<double> b(L);
                for (int k = 0; k < L; ++k) b[k] = s.taps_re[k];
                int d = grpdelay_mean_round(b);
                if (s.out_len > 0) {
                    d %= s.out_len;
                    if (d < 0) d += s.out_len;
                }
                sp.d.rot = d;
            } else if (s.align != RB200_ALIGN_DELAYED) {
                return fail(ctx, RB200_ERR_ARG, "set_waveform: FIR segments use DELAYED or GRPDELAY");
            }
        } else if (s.kind == RB200_SEG_MF_CIRC) {
            // one FFT tile of exactly out_len points keeps every lag, wrapped ones included: circular correlation
            if (s.align != RB200_ALIGN_LEADING_EDGE) return fail(ctx, RB200_ERR_ARG, "set_waveform: MF_CIRC segments use LEADING_EDGE");
            if ((s.out_len != 256 && s.out_len != 512 && s.out_len != 4096) || s.in_len > s.out_len || L > s.out_len)
                return fail(ctx, RB200_ERR_UNSUPPORTED, "set_waveform: MF_CIRC needs out_len in {256,512,4096} and in_len, n_taps <= out_len");
            corr_taps[i] = raw;
            sp.d.pre = 0;
            sp.d.rot = 0;
        } else {
            return fail(ctx, RB200_ERR_ARG, "set_waveform: unknown segment kind");
        }
        sp.nt = s.kind == RB200_SEG_MF_CIRC ? s.out_len : choose_nt(L, ctx->env.pc_nt);
        sp.d.t_off = (int)taps_all.size();
        for (int k = 0; k < L; ++k) {
            const cd t = corr_taps[i][k] * s.scale;    // direct kernel multiplies by conj(t) -> fold real scale
            taps_all.push_back(make_float2((float)t.real(), (float)t.imag()));
        }
        if (sp.nt) {
            sp.d.V = s.kind == RB200_SEG_MF_CIRC ? sp.nt : sp.nt - L + 1;
            // 256-sample tiles: an even number of valid lags keeps every tile's first sample on an even range cell, which the
            // tensor-map loads of pcw_kernel need (rows of two range cells); costs at most one lag per tile
            if (s.kind != RB200_SEG_MF_CIRC && sp.nt == 256 && sp.d.V > 2) sp.d.V &= ~1;
            sp.d.h_off = (int)h_total;
            h_total += sp.nt;
        }
        plan.segs.push_back(sp);
        plan.max_in_end = std::max(plan.max_in_end, s.in_start + s.in_len);
        plan.max_out_end = std::max(plan.max_out_end, s.out_start + s.out_len);
        plan.out_ranges.push_back({s.out_start, s.out_start + s.out_len});
    }
    std::sort(plan.out_ranges.begin(), plan.out_ranges.end());
    // spectra
    std::vector<float2> hperm(std::max<size_t>(h_total, 1));
    for (int i = 0; i < nseg; ++i) {
        const SegPlan& sp = plan.segs[i];
        if (!sp.nt) continue;
        const int NT = sp.nt;
        std::vector<cd> a(NT, cd(0, 0));
        for (int k = 0; k < sp.d.n_taps; ++k) a[k] = corr_taps[i][k];
        host_fft_pow2(a);
        const int R = (NT == 512) ? 8 : 16;
        int S = 0;
        for (int n = NT; n > 1; n /= R) ++S;
        for (int pos = 0; pos < NT; ++pos) {
            int f = 0, mul = 1, div = NT / R;
            for (int d = 0; d < S; ++d) {
                f += ((pos / div) % R) * mul;
                mul *= R;
                div /= R;
            }
            const cd h = std::conj(a[f]) * (segs[i].scale / (double)NT);
            hperm[sp.d.h_off + pos] = make_float2((float)h.real(), (float)h.imag());
        }
    }
    plan.h_entries = (int)h_total;
    CK(ctx, plan.hperm.ensure(hperm.size() * sizeof(float2)));
    CK(ctx, cudaMemcpyAsync(plan.hperm.p, hperm.data(), hperm.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, plan.taps.ensure(std::max<size_t>(taps_all.size(), 1) * sizeof(float2)));
    CK(ctx, cudaMemcpyAsync(plan.taps.p, taps_all.data(), taps_all.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    // tile lists and twiddles per class
    const int nts[3] = {256, 512, 4096};
    for (int ci = 0; ci < 3; ++ci) {
        std::vector<int2> tiles;
        for (int i = 0; i < nseg; ++i) {
            const SegPlan& sp = plan.segs[i];
            if (sp.nt != nts[ci] || sp.d.out_len == 0) continue;
            const int nt_tiles = (sp.d.out_len + sp.d.V - 1) / sp.d.V;
            for (int t = 0; t < nt_tiles; ++t) tiles.push_back(make_int2(i, t));
        }
        if (tiles.empty()) continue;
        plan.classes.emplace_back();
        ClassPlan& c = plan.classes.back();
        c.nt = nts[ci];
        c.n_tiles = (int)tiles.size();
        CK(ctx, c.tiles.ensure(tiles.size() * sizeof(int2)));
        CK(ctx, cudaMemcpyAsync(c.tiles.p, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
        std::vector<float2> tw;
        pc_build_twiddles(c.nt, tw);
        CK(ctx, c.tw.ensure(tw.size() * sizeof(float2)));
        CK(ctx, cudaMemcpyAsync(c.tw.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    plan.valid = true;
    return RB200_OK;
}

// run pulse compression: `in` wire int16 (n_groups = cpi*P groups of C lanes) or planar float2 (n_lines lines)
// does run_pc take the warp-private pcw_kernel for this plan / input?
static bool pcw_eligible(const rb200_ctx* ctx, const Plan& plan, bool wire, const void* in, int R, int C) {
    if (!wire || plan.classes.size() != 1 || plan.classes[0].nt != 256 || ctx->env.no_tma || ctx->env.no_pcw ||
        (reinterpret_cast<uintptr_t>(in) & 15) != 0)
        return false;
    PcParams q;
    memset(&q, 0, sizeof q);
    q.R = R;
    q.C = C;
    for (size_t i = 0; i < plan.segs.size(); ++i) {
        if (plan.segs[i].nt != 256) return false;
        q.segs[i] = plan.segs[i].d;
    }
    return pcw_plan_supported(q, (int)plan.segs.size(), plan.h_entries);
}

static int run_pc(rb200_ctx* ctx, const Plan& plan, bool wire, const void* in, float2* out, int R, int R_out,
                  int C, int P, int n_groups_wire, int n_lines, const float* gain, cudaStream_t st, int* cpi_done = nullptr,
                  int* started = nullptr, int pcw_sms = 0) {
    if (!plan.valid) return fail(ctx, RB200_ERR_NO_WAVEFORM, "no waveform plan: call rb200_set_waveform first");
    if (plan.max_in_end > R) return fail(ctx, RB200_ERR_INDEX, "waveform segment exceeds the PRT length (Index exceeds array bounds)");
    if (plan.max_out_end > R_out) return fail(ctx, RB200_ERR_INDEX, "waveform segment output exceeds the PRT length");
    const size_t out_lines = wire ? (size_t)n_groups_wire * C : (size_t)n_lines;
    // columns no segment writes stay zero (s_PC_0 = zeros(...))
    int cur = 0;
    for (auto& rg : plan.out_ranges) {
        if (rg.first > cur) { CK(ctx, launch_pc_zero_cols(out, out_lines, R_out, cur, rg.first, st)); ctx->launches++; }
        cur = std::max(cur, rg.second);
    }
    if (cur < R_out) { CK(ctx, launch_pc_zero_cols(out, out_lines, R_out, cur, R_out, st)); ctx->launches++; }
    PcParams p;
    memset(&p, 0, sizeof p);
    p.in = in;
    p.out = out;
    p.hperm = plan.hperm.as<float2>();
    p.gain = gain;
    for (size_t i = 0; i < plan.segs.size(); ++i) p.segs[i] = plan.segs[i].d;
    p.R = R;
    p.R_out = R_out;
    p.C = C;
    p.P = P;
    p.n_lines = n_lines;
    for (auto& c : plan.classes) {
        p.tw = c.tw.as<float2>();
        p.tiles = c.tiles.as<int2>();
        const int lt = pc_tile_lanes(c.nt, wire);
        const int n_groups = wire ? n_groups_wire : (n_lines + lt - 1) / lt;
        if (n_groups <= 0) continue;
        const bool pcw_ok = pcw_eligible(ctx, plan, wire, in, R, C);
        if (pcw_ok) {
            p.cpi_done = cpi_done;
            p.started = started;
            CK(ctx, launch_pcw(p, c.n_tiles, n_groups, pcw_sms > 0 ? pcw_sms : ctx->n_sms, plan.h_entries, ctx->env.coexist, st));
        }
        else if (wire && C == 16 && c.nt == 256 && plan.h_entries <= 2048 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && !ctx->env.no_tma)
            CK(ctx, launch_pc_fft_tma(p, c.n_tiles, n_groups, ctx->n_sms, ctx->pc_ctas_per_sm, plan.h_entries, st));
        else CK(ctx, launch_pc_fft(c.nt, wire, p, c.n_tiles, n_groups, st));
        ctx->launches++;
    }
    for (size_t i = 0; i < plan.segs.size(); ++i) {
        if (plan.segs[i].nt || plan.segs[i].d.out_len == 0) continue;
        const int lines = wire ? n_groups_wire * C : n_lines;
        if (lines <= 0) continue;
        CK(ctx, launch_pc_direct(wire, p, plan.taps.as<float2>(), (int)i, plan.segs[i].d.out_len, lines, st));
        ctx->launches++;
    }
    return RB200_OK;
}

static void factor_radices(int P, int* radix, int* n) {
    int k = 0, m = P;
    while (m % 8 == 0) { radix[k++] = 8; m /= 8; }
    while (m % 4 == 0) { radix[k++] = 4; m /= 4; }
    while (m % 2 == 0) { radix[k++] = 2; m /= 2; }
    for (int f = 3; m > 1 && k < 15; f += 2)
        while (m % f == 0 && k < 15) { radix[k++] = f; m /= f; }
    if (m > 1) radix[k++] = m;
    if (k == 0) radix[k++] = 1;
    *n = k;
}

static int get_mtd_plan(rb200_ctx* ctx, int P, double beta, MtdPlan** out) {
    long long bkey;
    memcpy(&bkey, &beta, sizeof bkey);
    auto key = std::make_pair(P, bkey);
    auto it = ctx->mtd_plans.find(key);
    if (it != ctx->mtd_plans.end()) { *out = it->second; return RB200_OK; }
    MtdPlan* mp = new MtdPlan();
    mp->P = P;
    mp->beta = beta;
    std::vector<double> w = kaiser_window(P, beta);
    std::vector<float> wf(P);
    for (int i = 0; i < P; ++i) wf[i] = (float)w[i];
    mp->h_window = wf;
    std::vector<float2> tw(P);
    for (int m = 0; m < P; ++m) {
        const double a = -2.0 * M_PI * m / P;
        tw[m] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    factor_radices(P, mp->radix, &mp->n_stages);
    cudaError_t e = mp->window.ensure(P * sizeof(float));
    if (e == cudaSuccess) e = mp->tw.ensure(P * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpyAsync(mp->window.p, wf.data(), P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mp->tw.p, tw.data(), P * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        delete mp;
        ctx->err = std::string("mtd plan: ") + cudaGetErrorString(e);
        return RB200_ERR_CUDA;
    }
    ctx->mtd_plans[key] = mp;
    *out = mp;
    return RB200_OK;
}

// MTD plan with a caller-supplied slow-time window of nw <= P entries (zero padding beyond), keyed by its contents
static int get_mtd_plan_custom(rb200_ctx* ctx, int P, const double* w, int nw, MtdPlan** out) {
    unsigned long long hsh = 1469598103934665603ull;
    const unsigned char* bytes = reinterpret_cast<const unsigned char*>(w);
    for (size_t i = 0; i < (size_t)nw * sizeof(double); ++i) hsh = (hsh ^ bytes[i]) * 1099511628211ull;
    hsh ^= (unsigned long long)nw * 0x9e3779b97f4a7c15ull;
    auto key = std::make_pair(-P, (long long)hsh);          // negative P: never collides with the Kaiser plans
    auto it = ctx->mtd_plans.find(key);
    if (it != ctx->mtd_plans.end()) { *out = it->second; return RB200_OK; }
    MtdPlan* mp = new MtdPlan();
    mp->P = P;
    std::vector<float> wf(P, 0.f);
    for (int i = 0; i < nw; ++i) wf[i] = (float)w[i];
    mp->h_window = wf;
    std::vector<float2> tw(P);
    for (int m = 0; m < P; ++m) {
        const double a = -2.0 * M_PI * m / P;
        tw[m] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    factor_radices(P, mp->radix, &mp->n_stages);
    cudaError_t e = mp->window.ensure(P * sizeof(float));
    if (e == cudaSuccess) e = mp->tw.ensure(P * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpyAsync(mp->window.p, wf.data(), P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mp->tw.p, tw.data(), P * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        delete mp;
        ctx->err = std::string("mtd plan: ") + cudaGetErrorString(e);
        return RB200_ERR_CUDA;
    }
    ctx->mtd_plans[key] = mp;
    *out = mp;
    return RB200_OK;
}

// zero-velocity rows, 0-based inclusive; lo > hi when disabled.  MP/fun_0v_pressing.m:4-6.
static int zero_v_rows(int P, int div, int* lo, int* hi) {
    if (div <= 0) { *lo = 1; *hi = 0; return RB200_OK; }
    const long z = mround(P / 2.0), h = mround((double)P / div);
    const long a = z - h, b = z + h;            // 1-based inclusive
    if (a < 1 || b > P) return RB200_ERR_INDEX;
    *lo = (int)a - 1;
    *hi = (int)b - 1;
    return RB200_OK;
}

struct FusedV {            // optional velocity-CFAR fusion into the shared-memory MTD kernels
    const CfarParams* cf;
    float t_v;
    void* dets;
    int* det_count;
    uint32_t* vmask;
    int* err_flag;
};

static int run_mtd(rb200_ctx* ctx, const float2* in, float* out, int P, int in_ld, int out_ld, int cols, int n_slabs,
                   double beta, int zero_div, int mti_lag, cudaStream_t st, const FusedV* fv = nullptr, float2* out_c = nullptr,
                   int crop_lo = 0, int crop_hi = -1) {
    if (P < 1) return fail(ctx, RB200_ERR_ARG, "MTD: P < 1");
    if (!mtd_has_fast_path(P) && P > mtd_generic_max_p()) return fail(ctx, RB200_ERR_UNSUPPORTED, "MTD: P beyond the generic kernel's shared-memory envelope (12288)");
    MtdPlan* mp = nullptr;
    int rc = get_mtd_plan(ctx, P, beta, &mp);
    if (rc) return rc;
    MtdParams p;
    memset(&p, 0, sizeof p);
    p.in = in;
    p.out = out;
    p.window = mp->window.as<float>();
    p.tw = mp->tw.as<float2>();
    p.P = P;
    p.in_ld = in_ld;
    p.out_ld = out_ld;
    p.cols = cols;
    p.mti_lag = mti_lag;
    rc = zero_v_rows(P, zero_div, &p.zv_lo, &p.zv_hi);
    if (rc) return fail(ctx, rc, "fun_0v_pressing: Index in position 1 is invalid");
    p.n_stages = mp->n_stages;
    for (int i = 0; i < mp->n_stages; ++i) p.radix[i] = mp->radix[i];
    p.out_c = out_c;
    p.crop_lo = crop_lo;
    p.crop_hi = crop_hi;
    p.n_sms = ctx->n_sms;
    p.no_tma = ctx->env.no_tma_mtd ? 1 : 0;
    if (ctx->env.mtd_tc && P == 64 && mti_lag == 0 && !fv && !out_c) {
        // experiment: DFT-by-GEMM on tcgen05 (window, fftshift and the zero-velocity rows are folded into the matrix)
        if (mp->tc_zlo != p.zv_lo || mp->tc_zhi != p.zv_hi || !mp->tc_mat.p) {
            std::vector<uint16_t> a;
            mtd64_tc_build_matrix(mp->h_window.data(), p.zv_lo, p.zv_hi, a);
            CK(ctx, mp->tc_mat.ensure(mtd64_tc_matrix_bytes()));
            CK(ctx, cudaMemcpyAsync(mp->tc_mat.p, a.data(), mtd64_tc_matrix_bytes(), cudaMemcpyHostToDevice, st));
            CK(ctx, cudaStreamSynchronize(st));
            mp->tc_zlo = p.zv_lo;
            mp->tc_zhi = p.zv_hi;
        }
        CK(ctx, launch_mtd64_tc(in, out, mp->tc_mat.p, in_ld, out_ld, cols, n_slabs, ctx->n_sms, st));
        ctx->launches++;
        return RB200_OK;
    }
    if (fv) {
        p.cfar_on = 1;
        p.cf = *fv->cf;
        p.t_v = fv->t_v;
        p.dets = fv->dets;
        p.det_count = fv->det_count;
        p.vmask = fv->vmask;
        p.err_flag = fv->err_flag;
    }
    CK(ctx, launch_mtd(p, n_slabs, st));
    ctx->launches++;
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------------
extern "C" int rb200_version(void) { return RB200_VERSION_MAJOR * 100 + RB200_VERSION_MINOR; }

static void default_config(rb200_config* c) {
    memset(c, 0, sizeof *c);
    c->struct_size = (int32_t)sizeof(rb200_config);
    c->n_prt = 64;
    c->n_range = 4096;
    c->n_lanes = 16;
    c->max_cpi = 1;
    c->max_det = 65536;
    c->zero_v_div = 150;
    c->kaiser_beta = 8.0;
    c->cfar_ref_r = c->cfar_ref_v = 5;
    c->cfar_guard_r = c->cfar_guard_v = 7;
    c->cfar_t_r = c->cfar_t_v = 5.0;
    c->cfar_range_stage = 1;
}

static int validate_cfar(rb200_ctx* ctx, const rb200_config& c) {
    if (c.cfar_ref_r < 1 || c.cfar_ref_v < 1 || c.cfar_guard_r < 0 || c.cfar_guard_v < 0)
        return fail(ctx, RB200_ERR_ARG, "config: CFAR ref cells must be >= 1 and guard cells >= 0");
    if ((c.cfar_method_r != 0 && c.cfar_method_r != 1) || (c.cfar_method_v != 0 && c.cfar_method_v != 1))
        return fail(ctx, RB200_ERR_ARG, "config: CFAR method must be 0 (GO) or 1 (SO)");
    return RB200_OK;
}

extern "C" int rb200_create(rb200_ctx** out, int device, const rb200_config* cfg) {
    if (!out) { g_create_error = "rb200_create: out is NULL"; return RB200_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("rb200_create: no CUDA device (") + cudaGetErrorString(e) + "); libradar_b200 has no CPU fallback";
        return RB200_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "rb200_create: device index out of range"; return RB200_ERR_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return RB200_ERR_CUDA; }
    rb200_ctx* c = new rb200_ctx();
    c->device = device;
    default_config(&c->cfg);
    if (cfg) {
        if (cfg->struct_size != (int32_t)sizeof(rb200_config)) { delete c; g_create_error = "rb200_create: rb200_config.struct_size mismatch"; return RB200_ERR_ARG; }
        c->cfg = *cfg;
    }
    const rb200_config& k = c->cfg;
    if (k.n_prt < 1 || k.n_range < 1 || k.n_lanes < 1 || k.n_lanes > 255 || k.max_cpi < 1 || k.max_det < 1 || k.n_prt > 65535) {
        delete c;
        g_create_error = "rb200_create: geometry out of range (n_prt 1..65535, n_lanes 1..255, sizes >= 1)";
        return RB200_ERR_ARG;
    }
    if (validate_cfar(c, k)) { g_create_error = c->err; delete c; return RB200_ERR_ARG; }
    cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device);
    c->env.read();
    cudaDeviceGetAttribute(&c->coop_launch, cudaDevAttrCooperativeLaunch, device);
    if (const char* e1 = getenv("RB200_PC_CTAS")) c->pc_ctas_per_sm = atoi(e1);
    if (const char* e2 = getenv("RB200_MTD_CTAS")) c->mtd_ctas_per_sm = atoi(e2);
    else if (c->env.coexist) c->mtd_ctas_per_sm = 1;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = c->counters.ensure(4 * sizeof(int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming);
    for (int i = 0; i < rb200_ctx::kMaxSlots && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&c->slots[i].stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->slots[i].done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = c->slots[i].count.ensure(2 * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(c->slots[i].count.p, 0, 2 * sizeof(int));
    }
    if (e == cudaSuccess) e = c->errflag.ensure(sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_counts, 4 * sizeof(int));
    if (e != cudaSuccess) {
        g_create_error = std::string("rb200_create: ") + cudaGetErrorString(e);
        rb200_destroy(c);
        return RB200_ERR_CUDA;
    }
    *out = c;
    return RB200_OK;
}

extern "C" int rb200_destroy(rb200_ctx* c) {
    if (!c) return RB200_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->plan.release();
    c->dmx_plan.release();
    c->pcz_plan.release();
    for (auto& kv : c->mtd_plans) {
        kv.second->window.release();
        kv.second->tw.release();
        delete kv.second;
    }
    DevBuf* bufs[] = {&c->gain, &c->raw, &c->pc, &c->rdm, &c->dets_v, &c->dets_2d, &c->counters, &c->vmask, &c->errflag, &c->colmask, &c->ring, &c->megactr, &c->dbf_w,
                      &c->s_in_re, &c->s_in_im, &c->s_a, &c->s_b, &c->s_c, &c->s_out_re, &c->s_out_im, &c->s_u8a, &c->s_u8b, &c->s_idx};
    for (DevBuf* b : bufs) b->release();
    if (c->h_dets) cudaFreeHost(c->h_dets);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    for (int i = 0; i < rb200_ctx::kMaxSlots; ++i) {
        rb200_ctx::Slot& sl = c->slots[i];
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        sl.pc.release(); sl.colmask.release(); sl.vlist.release(); sl.count.release(); sl.raw.release(); sl.rdm.release(); sl.beams.release();
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    for (cudaEvent_t e : c->stage_events) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return RB200_OK;
}

// ---- process-wide shared contexts (one per device), reference counted --------------------------------------------
namespace {
struct SharedCtx { rb200_ctx* ctx = nullptr; int refs = 0; };
SharedCtx g_shared[64];
}
extern "C" int rb200_shared_context_acquire(rb200_ctx** out, int device) {
    if (!out || device < 0 || device >= 64) { g_create_error = "rb200_shared_context_acquire: bad argument"; return RB200_ERR_ARG; }
    SharedCtx& s = g_shared[device];
    if (!s.ctx) {
        int rc = rb200_create(&s.ctx, device, nullptr);
        if (rc) { s.ctx = nullptr; *out = nullptr; return rc; }
        s.refs = 0;
    }
    ++s.refs;
    *out = s.ctx;
    return RB200_OK;
}
extern "C" int rb200_shared_context_release(int device) {
    if (device < 0 || device >= 64) return RB200_ERR_ARG;
    SharedCtx& s = g_shared[device];
    if (!s.ctx || s.refs <= 0) return RB200_ERR_ARG;
    if (--s.refs == 0) {
        rb200_destroy(s.ctx);
        s.ctx = nullptr;
    }
    return RB200_OK;
}
extern "C" int rb200_set_plan_tag(rb200_ctx* c, uint64_t tag) {
    if (!c) return RB200_ERR_ARG;
    c->plan_tag = tag;
    return RB200_OK;
}
extern "C" uint64_t rb200_get_plan_tag(const rb200_ctx* c) { return c ? c->plan_tag : 0; }

extern "C" const char* rb200_last_error(const rb200_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

extern "C" int rb200_get_config(const rb200_ctx* c, rb200_config* out) {
    if (!c || !out) return RB200_ERR_ARG;
    *out = c->cfg;
    return RB200_OK;
}

extern "C" int rb200_set_cfar(rb200_ctx* c, const rb200_config* cfg) {
    if (!c || !cfg) return RB200_ERR_ARG;
    rb200_config n = c->cfg;
    n.cfar_ref_r = cfg->cfar_ref_r; n.cfar_guard_r = cfg->cfar_guard_r; n.cfar_method_r = cfg->cfar_method_r;
    n.cfar_ref_v = cfg->cfar_ref_v; n.cfar_guard_v = cfg->cfar_guard_v; n.cfar_method_v = cfg->cfar_method_v;
    n.cfar_t_r = cfg->cfar_t_r; n.cfar_t_v = cfg->cfar_t_v; n.cfar_n0 = cfg->cfar_n0; n.cfar_range_stage = cfg->cfar_range_stage;
    int rc = validate_cfar(c, n);
    if (rc) return rc;
    c->cfg = n;
    return RB200_OK;
}

extern "C" int rb200_set_waveform(rb200_ctx* c, const rb200_segment* segs, int nseg) {
    if (!c) return RB200_ERR_ARG;
    cudaSetDevice(c->device);
    c->plan_tag = 0;
    return build_plan(c, c->plan, segs, nseg);
}

extern "C" int rb200_set_stc(rb200_ctx* c, const double* stc_db, int n) {
    if (!c || n < 0 || (n > 0 && !stc_db)) return RB200_ERR_ARG;
    cudaSetDevice(c->device);
    if (n == 0) { c->gain_n = 0; return RB200_OK; }
    const int R = c->cfg.n_range;
    if (n > R) return fail(c, RB200_ERR_DIM_MISMATCH, "fun_iSTC: STC curve longer than the PRT (arrays have incompatible sizes)");
    std::vector<float> g(R, 1.0f);
    for (int i = 0; i < n; ++i) g[i] = (float)std::pow(10.0, stc_db[i] / 20.0);
    CK(c, c->gain.ensure(R * sizeof(float)));
    CK(c, cudaMemcpyAsync(c->gain.p, g.data(), R * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    c->gain_n = n;
    return RB200_OK;
}

extern "C" int rb200_set_cfar_segments(rb200_ctx* c, const int32_t* lo, const int32_t* hi, int n) {
    if (!c || n < 0 || n > 4 || (n > 0 && (!lo || !hi))) return fail(c, RB200_ERR_ARG, "set_cfar_segments: 0..4 segments");
    CfarSegs s = {};
    for (int i = 0; i < n; ++i) {
        if (lo[i] < 0 || hi[i] <= lo[i] || (i > 0 && lo[i] < hi[i - 1])) return fail(c, RB200_ERR_ARG, "set_cfar_segments: segments must be ascending, non-empty and disjoint");
        s.lo[i] = lo[i];
        s.hi[i] = hi[i];
    }
    s.n = n;
    c->cfar_segs = s;
    return RB200_OK;
}

extern "C" int rb200_set_dbf(rb200_ctx* c, const double* w_re, const double* w_im, int n_beams) {
    if (!c || n_beams < 0 || n_beams > 255 || (n_beams > 0 && !w_re)) return fail(c, RB200_ERR_ARG, "set_dbf: bad argument");
    cudaSetDevice(c->device);
    if (n_beams == 0) { c->dbf_beams = 0; return RB200_OK; }
    const int nch = c->cfg.n_lanes;
    if ((size_t)n_beams * nch * sizeof(float2) > 48 * 1024) return fail(c, RB200_ERR_UNSUPPORTED, "set_dbf: weight matrix larger than 48 KB");
    // DBF_coeffs_data_C is n_beams x n_channels, column-major; device copy is [beam][channel]
    std::vector<float2> w((size_t)n_beams * nch);
    for (int b = 0; b < n_beams; ++b)
        for (int ch = 0; ch < nch; ++ch) {
            const size_t k = (size_t)b + (size_t)n_beams * ch;
            w[(size_t)b * nch + ch] = make_float2((float)w_re[k], w_im ? (float)w_im[k] : 0.f);
        }
    CK(c, c->dbf_w.ensure(w.size() * sizeof(float2)));
    CK(c, cudaMemcpyAsync(c->dbf_w.p, w.data(), w.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    c->dbf_beams = n_beams;
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------
// MATLAB-layout entry points
// ---------------------------------------------------------------------------------------------
static int upload_z(rb200_ctx* c, const double* re, const double* im, size_t n, const double** dre, const double** dim) {
    CK(c, c->s_in_re.ensure(std::max<size_t>(n, 1) * sizeof(double)));
    CK(c, cudaMemcpyAsync(c->s_in_re.p, re, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    *dre = c->s_in_re.as<double>();
    *dim = nullptr;
    if (im) {
        CK(c, c->s_in_im.ensure(std::max<size_t>(n, 1) * sizeof(double)));
        CK(c, cudaMemcpyAsync(c->s_in_im.p, im, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        *dim = c->s_in_im.as<double>();
    }
    return RB200_OK;
}

extern "C" int rb200_pulse_compression_z(rb200_ctx* c, const double* s0_re, const double* s0_im, int L,
                                         const double* echo_re, const double* echo_im, int M, double* out_re, double* out_im) {
    if (!c || !s0_re || L < 1 || M < 0 || (M > 0 && !echo_re) || !out_re || !out_im) return fail(c, RB200_ERR_ARG, "pulse_compression: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    const int N = L + M - 1;
    if (N < 1) return RB200_OK;
    if (M == 0) {   // fft of an empty vector padded to N is zero -> output zeros(1,N); device memset, no CPU maths
        CK(c, c->s_out_re.ensure(N * sizeof(double)));
        CK(c, cudaMemsetAsync(c->s_out_re.p, 0, N * sizeof(double), c->stream));
        CK(c, cudaMemcpyAsync(out_re, c->s_out_re.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemcpyAsync(out_im, c->s_out_re.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        return RB200_OK;
    }
    // full convolution with conj(flip(s0)): y[m] = sum_j x[m-(L-1)+j] conj(s0[j])  (MP/fun_pulse_compression.m:4,19-22)
    // The plan (reference spectrum, tiles, twiddles) is cached on (L, M, taps): a caller looping over PRTs with one pulse --
    // MP/fun_lss_pulse_compression.m:24-37 does exactly that -- pays for it once, not per call like the M-code (:20).
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&](const void* ptr, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(ptr);
        for (size_t i = 0; i < n; ++i) { key ^= b[i]; key *= 1099511628211ull; }
    };
    mix(&L, sizeof L);
    mix(&M, sizeof M);
    mix(s0_re, (size_t)L * sizeof(double));
    if (s0_im) mix(s0_im, (size_t)L * sizeof(double));
    else { const int z = 0; mix(&z, sizeof z); }
    Plan& plan = c->pcz_plan;
    if (!plan.valid || c->pcz_key != key) {
        rb200_segment s;
        memset(&s, 0, sizeof s);
        s.in_start = 0; s.in_len = M; s.out_start = 0; s.out_len = N;
        s.kind = RB200_SEG_MF; s.align = RB200_ALIGN_LEADING_EDGE;
        s.n_taps = L; s.taps_re = s0_re; s.taps_im = s0_im; s.scale = 1.0;
        int rc = build_plan(c, plan, &s, 1);
        if (rc) { plan.release(); return rc; }
        plan.segs[0].d.pre = L - 1;
        c->pcz_key = key;
    }
    const double *dre, *dim;
    int rc = upload_z(c, echo_re, echo_im, M, &dre, &dim);
    if (rc) return rc;
    CK(c, c->s_a.ensure((size_t)M * sizeof(float2)));
    CK(c, c->s_b.ensure((size_t)N * sizeof(float2)));
    CK(c, c->s_out_re.ensure(N * sizeof(double)));
    CK(c, c->s_out_im.ensure(N * sizeof(double)));
    CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>(), 1, M, c->stream));
    c->launches++;
    rc = run_pc(c, plan, false, c->s_a.p, c->s_b.as<float2>(), M, N, 1, 1, 0, 1, nullptr, c->stream);
    if (rc) return rc;
    CK(c, launch_planar_to_z(c->s_b.as<float2>(), c->s_out_re.as<double>(), c->s_out_im.as<double>(), 1, N, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out_re, c->s_out_re.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(out_im, c->s_out_im.p, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_lss_pulse_compression_z(rb200_ctx* c, const double* echo_re, const double* echo_im, int P, int R,
                                             double* out_re, double* out_im) {
    if (!c || !echo_re || P < 1 || R < 1 || !out_re || !out_im) return fail(c, RB200_ERR_ARG, "lss_pulse_compression: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    const size_t n = (size_t)P * R;
    const double *dre, *dim;
    int rc = upload_z(c, echo_re, echo_im, n, &dre, &dim);
    if (rc) return rc;
    CK(c, c->s_a.ensure(n * sizeof(float2)));
    CK(c, c->s_b.ensure(n * sizeof(float2)));
    CK(c, c->s_out_re.ensure(n * sizeof(double)));
    CK(c, c->s_out_im.ensure(n * sizeof(double)));
    CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>(), P, R, c->stream));
    c->launches++;
    rc = run_pc(c, c->plan, false, c->s_a.p, c->s_b.as<float2>(), R, R, 1, 1, 0, P, nullptr, c->stream);
    if (rc) return rc;
    CK(c, launch_planar_to_z(c->s_b.as<float2>(), c->s_out_re.as<double>(), c->s_out_im.as<double>(), P, R, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out_re, c->s_out_re.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(out_im, c->s_out_im.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_process_mtd_z(rb200_ctx* c, const double* re, const double* im, int rows, int cols, int len_prt, int num_prt,
                                   double beta, double* out) {
    if (!c || !re || rows < 1 || cols < 0 || len_prt < 0 || num_prt < 1 || !out) return fail(c, RB200_ERR_ARG, "process_mtd: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    if (rows != num_prt) {
        if (rows == 1 || num_prt == 1) return fail(c, RB200_ERR_UNSUPPORTED, "process_mtd: implicit expansion (rows==1 or Num_PRTperFrame==1) is not supported");
        return fail(c, RB200_ERR_DIM_MISMATCH, "fun_Process_MTD: arrays have incompatible sizes (rows of ProSignal != Num_PRTperFrame)");
    }
    if (len_prt > cols) return fail(c, RB200_ERR_INDEX, "fun_Process_MTD: Index in position 2 exceeds array bounds (Len_PRT > columns)");
    if (len_prt == 0) return RB200_OK;
    const int P = rows;
    const size_t n = (size_t)P * cols, no = (size_t)P * len_prt;
    const double *dre, *dim;
    int rc = upload_z(c, re, im, n, &dre, &dim);
    if (rc) return rc;
    CK(c, c->s_a.ensure(n * sizeof(float2)));
    CK(c, c->s_c.ensure(no * sizeof(float)));
    CK(c, c->s_out_re.ensure(no * sizeof(double)));
    CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>(), P, cols, c->stream));
    c->launches++;
    rc = run_mtd(c, c->s_a.as<float2>(), c->s_c.as<float>(), P, cols, len_prt, len_prt, 1, beta, 0, 0, c->stream);
    if (rc) return rc;
    CK(c, launch_f32_rowmajor_to_d_colmajor(c->s_c.as<float>(), c->s_out_re.as<double>(), P, len_prt, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out, c->s_out_re.p, no * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_zero_v_pressing_d(rb200_ctx* c, const double* mtd, int P, int R, int div, double* out) {
    if (!c || !mtd || !out || P < 1 || R < 0 || div < 1) return fail(c, RB200_ERR_ARG, "zero_v_pressing: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    int lo, hi;
    if (zero_v_rows(P, div, &lo, &hi)) return fail(c, RB200_ERR_INDEX, "fun_0v_pressing: Index in position 1 is invalid");
    const size_t n = (size_t)P * R;
    if (n == 0) return RB200_OK;
    CK(c, c->s_in_re.ensure(n * sizeof(double)));
    CK(c, c->s_out_re.ensure(n * sizeof(double)));
    CK(c, cudaMemcpyAsync(c->s_in_re.p, mtd, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, launch_zero_rows_d_colmajor(c->s_in_re.as<double>(), c->s_out_re.as<double>(), P, R, lo, hi, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out, c->s_out_re.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

// interleaved = false: echo_re / echo_im are split column-major arrays (mxGetPr / mxGetPi); true: echo_re points at interleaved
// complex doubles (MATLAB -R2018a mxGetComplexDoubles, numpy complex128), echo_im is ignored
static int upload_echo(rb200_ctx* c, const double* echo_re, const double* echo_im, bool interleaved, size_t n, const double** dre,
                       const double** dim, int* es) {
    if (!interleaved) { *es = 1; return upload_z(c, echo_re, echo_im, n, dre, dim); }
    CK(c, c->s_in_re.ensure(std::max<size_t>(n, 1) * 2 * sizeof(double)));
    CK(c, cudaMemcpyAsync(c->s_in_re.p, echo_re, n * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    *dre = c->s_in_re.as<double>();
    *dim = c->s_in_re.as<double>() + 1;
    *es = 2;
    return RB200_OK;
}

static int mtd_produce_impl(rb200_ctx* c, const double* echo_re, const double* echo_im, bool interleaved, int P, int R, double beta,
                            int zero_v_div, double* out) {
    if (!c || !echo_re || P < 1 || R < 1 || !out) return fail(c, RB200_ERR_ARG, "mtd_produce: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    const size_t n = (size_t)P * R;
    const double *dre, *dim;
    int es;
    int rc = upload_echo(c, echo_re, echo_im, interleaved, n, &dre, &dim, &es);
    if (rc) return rc;
    CK(c, c->s_a.ensure(n * sizeof(float2)));
    CK(c, c->s_b.ensure(n * sizeof(float2)));
    CK(c, c->s_c.ensure(n * sizeof(float)));
    CK(c, c->s_out_re.ensure(n * sizeof(double)));
    CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>(), P, R, c->stream, es));
    c->launches++;
    rc = run_pc(c, c->plan, false, c->s_a.p, c->s_b.as<float2>(), R, R, 1, 1, 0, P, nullptr, c->stream);
    if (rc) return rc;
    rc = run_mtd(c, c->s_b.as<float2>(), c->s_c.as<float>(), P, R, R, R, 1, beta, zero_v_div, c->cfg.mti_lag, c->stream);
    if (rc) return rc;
    CK(c, launch_f32_rowmajor_to_d_colmajor(c->s_c.as<float>(), c->s_out_re.as<double>(), P, R, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out, c->s_out_re.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_mtd_produce_z(rb200_ctx* c, const double* echo_re, const double* echo_im, int P, int R, double beta,
                                   int zero_v_div, double* out) {
    return mtd_produce_impl(c, echo_re, echo_im, false, P, R, beta, zero_v_div, out);
}
extern "C" int rb200_mtd_produce_c(rb200_ctx* c, const double* echo_ri, int P, int R, double beta, int zero_v_div, double* out) {
    return mtd_produce_impl(c, echo_ri, nullptr, true, P, R, beta, zero_v_div, out);
}

// Crop-aware fun_MTD_produce: the caller keeps rows row_lo..row_hi of the result (MP/main_produce_dataset_win_xzr.m:39-40 keeps
// 691:845 of 1536).  Pulse compression (along range, the same operator for every PRT) and the windowed slow-time transform
// (along PRT, the same operator for every range cell) commute, so the transform runs FIRST and only the kept Doppler rows
// are pulse-compressed, converted and returned: ~10x less PC work and ~10x fewer result bytes for the reference's crop.
static int mtd_produce_rows_impl(rb200_ctx* c, const double* echo_re, const double* echo_im, bool interleaved, int P, int R, double beta,
                                 int zero_v_div, int row_lo, int row_hi, double* out) {
    if (!c || !echo_re || P < 1 || R < 1 || !out) return fail(c, RB200_ERR_ARG, "mtd_produce_rows: bad argument");
    if (row_lo < 1 || row_hi > P || row_lo > row_hi)
        return fail(c, RB200_ERR_INDEX, "fun_MTD_produce: Index in position 1 exceeds array bounds (row crop outside 1..P)");
    if (!mtd_has_fast_path(P) && P > mtd_generic_max_p()) return fail(c, RB200_ERR_UNSUPPORTED, "MTD: P beyond the generic kernel's shared-memory envelope (12288)");
    cudaSetDevice(c->device);
    c->launches = 0;
    const size_t n = (size_t)P * R;
    const int nrow = row_hi - row_lo + 1;
    const size_t nc = (size_t)nrow * R;
    const double *dre, *dim;
    int es;
    int rc = upload_echo(c, echo_re, echo_im, interleaved, n, &dre, &dim, &es);
    if (rc) return rc;
    CK(c, c->s_a.ensure(n * sizeof(float2)));
    CK(c, c->s_b.ensure(nc * sizeof(float2)));
    CK(c, c->s_out_re.ensure(nc * sizeof(double)));
    CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>(), P, R, c->stream, es));
    c->launches++;
    rc = run_mtd(c, c->s_a.as<float2>(), nullptr, P, R, R, R, 1, beta, zero_v_div, c->cfg.mti_lag, c->stream, nullptr, c->s_b.as<float2>(),
                 row_lo - 1, row_hi - 1);
    if (rc) return rc;
    rc = run_pc(c, c->plan, false, c->s_b.p, c->s_a.as<float2>(), R, R, 1, 1, 0, nrow, nullptr, c->stream);      // kept rows only
    if (rc) return rc;
    CK(c, launch_abs_planar_to_d_colmajor(c->s_a.as<float2>(), c->s_out_re.as<double>(), nrow, R, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out, c->s_out_re.p, nc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_mtd_produce_rows_z(rb200_ctx* c, const double* echo_re, const double* echo_im, int P, int R, double beta,
                                        int zero_v_div, int row_lo, int row_hi, double* out) {
    return mtd_produce_rows_impl(c, echo_re, echo_im, false, P, R, beta, zero_v_div, row_lo, row_hi, out);
}
extern "C" int rb200_mtd_produce_rows_c(rb200_ctx* c, const double* echo_ri, int P, int R, double beta, int zero_v_div, int row_lo,
                                        int row_hi, double* out) {
    return mtd_produce_rows_impl(c, echo_ri, nullptr, true, P, R, beta, zero_v_div, row_lo, row_hi, out);
}

extern "C" int rb200_mtd_produce_windows_z(rb200_ctx* c, const double* echo_re, const double* echo_im, int P_total, int R, int win_len,
                                           const int32_t* row_start, int n_win, double beta, int zero_v_div, double* out) {
    if (!c || !echo_re || !row_start || !out || P_total < 1 || R < 1 || win_len < 1 || n_win < 1)
        return fail(c, RB200_ERR_ARG, "mtd_produce_windows: bad argument");
    for (int i = 0; i < n_win; ++i)
        if (row_start[i] < 0 || row_start[i] + win_len > P_total)
            return fail(c, RB200_ERR_INDEX, "main_produce_dataset_win: Index in position 1 exceeds array bounds (window past the last PRT)");
    cudaSetDevice(c->device);
    c->launches = 0;
    const size_t n = (size_t)P_total * R, nw = (size_t)win_len * R;
    const double *dre, *dim;
    int rc = upload_z(c, echo_re, echo_im, n, &dre, &dim);
    if (rc) return rc;
    CK(c, c->s_a.ensure(n * sizeof(float2)));
    CK(c, c->s_b.ensure(n * sizeof(float2)));
    CK(c, c->s_c.ensure(nw * sizeof(float)));
    CK(c, c->s_out_re.ensure(nw * sizeof(double)));
    CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>(), P_total, R, c->stream));
    c->launches++;
    rc = run_pc(c, c->plan, false, c->s_a.p, c->s_b.as<float2>(), R, R, 1, 1, 0, P_total, nullptr, c->stream);   // once for all windows
    if (rc) return rc;
    for (int i = 0; i < n_win; ++i) {
        rc = run_mtd(c, c->s_b.as<float2>() + (size_t)row_start[i] * R, c->s_c.as<float>(), win_len, R, R, R, 1, beta, zero_v_div,
                     c->cfg.mti_lag, c->stream);
        if (rc) return rc;
        CK(c, launch_f32_rowmajor_to_d_colmajor(c->s_c.as<float>(), c->s_out_re.as<double>(), win_len, R, c->stream));
        c->launches++;
        CK(c, cudaMemcpyAsync(out + (size_t)i * nw, c->s_out_re.p, nw * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_dmx_process_z(rb200_ctx* c, const double* left_re, const double* left_im, const double* right_re,
                                   const double* right_im, int P, int n_range, int n_short, const double* fir_taps, int n_fir,
                                   const double* mf_re, const double* mf_im, int n_mf, int fft_num, const double* mtd_window,
                                   int mtd_fft_num, int n_blank, double* sum_short, double* diff_short, double* sum_long,
                                   double* diff_long) {
    if (!c || !left_re || !right_re || !mf_re || !mtd_window || P < 1 || n_range < 1 || n_short < 0 || n_short >= n_range || n_mf < 1 ||
        (n_short > 0 && (!fir_taps || n_fir < 1)) || (n_short == 0 && (sum_short || diff_short)))
        return fail(c, RB200_ERR_ARG, "dmx_process: bad argument");
    const int n_long = n_range - n_short;
    if ((fft_num != 256 && fft_num != 512 && fft_num != 4096) || n_long > fft_num || n_mf > fft_num)
        return fail(c, RB200_ERR_UNSUPPORTED, "dmx_process: FFT_num must be 256, 512 or 4096 and hold the long-pulse samples and the reference");
    if (mtd_fft_num < P || mtd_fft_num > mtd_generic_max_p())
        return fail(c, RB200_ERR_UNSUPPORTED, "dmx_process: mtd_FFT_num must be in [prtNum, 12288]");
    if (n_blank >= 0 && 2 * n_blank + 1 > mtd_fft_num) return fail(c, RB200_ERR_INDEX, "dmx_process: zeroSetFlagMTD exceeds the Doppler axis");
    cudaSetDevice(c->device);
    c->launches = 0;
    // ---- waveform plan (rebuilt only when the taps or the geometry change) ----
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&key](const void* p, size_t nbytes) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < nbytes; ++i) key = (key ^ b[i]) * 1099511628211ull;
    };
    const int geo[5] = {n_range, n_short, n_fir, n_mf, fft_num};
    mix(geo, sizeof geo);
    if (n_short > 0) mix(fir_taps, (size_t)n_fir * sizeof(double));
    mix(mf_re, (size_t)n_mf * sizeof(double));
    if (mf_im) mix(mf_im, (size_t)n_mf * sizeof(double));
    if (!c->dmx_plan.valid || c->dmx_key != key) {
        rb200_segment segs[2];
        memset(segs, 0, sizeof segs);
        int ns = 0;
        if (n_short > 0) {
            segs[ns].kind = RB200_SEG_FIR;
            segs[ns].align = RB200_ALIGN_DELAYED;
            segs[ns].in_start = 0;
            segs[ns].in_len = n_short;
            segs[ns].out_start = 0;
            segs[ns].out_len = n_short;
            segs[ns].taps_re = fir_taps;
            segs[ns].n_taps = n_fir;
            segs[ns].scale = 1.0;
            ++ns;
        }
        segs[ns].kind = RB200_SEG_MF_CIRC;
        segs[ns].align = RB200_ALIGN_LEADING_EDGE;
        segs[ns].in_start = n_short;
        segs[ns].in_len = n_long;
        segs[ns].out_start = n_short;
        segs[ns].out_len = fft_num;
        segs[ns].taps_re = mf_re;
        segs[ns].taps_im = mf_im;
        segs[ns].n_taps = n_mf;
        segs[ns].scale = 1.0;
        ++ns;
        int rc = build_plan(c, c->dmx_plan, segs, ns);
        if (rc) return rc;
        c->dmx_key = key;
    }
    // ---- both beams as planar lines [beam][prt][range] ----
    const int R_out = n_short + fft_num;
    const size_t n_in = (size_t)P * n_range, n_pc = (size_t)P * R_out, n_rdm = (size_t)mtd_fft_num * R_out;
    CK(c, c->s_a.ensure(2 * n_in * sizeof(float2)));
    CK(c, c->s_b.ensure(2 * n_pc * sizeof(float2)));
    CK(c, c->s_c.ensure(2 * n_rdm * sizeof(float)));
    const double* beams[2][2] = {{left_re, left_im}, {right_re, right_im}};
    for (int b = 0; b < 2; ++b) {
        const double *dre, *dim;
        int rc = upload_z(c, beams[b][0], beams[b][1], n_in, &dre, &dim);
        if (rc) return rc;
        CK(c, launch_z_to_planar(dre, dim, c->s_a.as<float2>() + b * n_in, P, n_range, c->stream));
        c->launches++;
        CK(c, cudaStreamSynchronize(c->stream));     // the staging buffers are re-used by the second beam
    }
    int rc = run_pc(c, c->dmx_plan, false, c->s_a.p, c->s_b.as<float2>(), n_range, R_out, 1, 1, 0, 2 * P, nullptr, c->stream);
    if (rc) return rc;
    // ---- slow-time FFT, zero-padded to mtd_fft_num, natural order ----
    MtdPlan* mp = nullptr;
    rc = get_mtd_plan_custom(c, mtd_fft_num, mtd_window, P, &mp);
    if (rc) return rc;
    MtdParams p;
    memset(&p, 0, sizeof p);
    p.in = c->s_b.as<float2>();
    p.out = c->s_c.as<float>();
    p.window = mp->window.as<float>();
    p.tw = mp->tw.as<float2>();
    p.P = mtd_fft_num;
    p.in_rows = P;
    p.no_shift = 1;
    p.in_ld = R_out;
    p.out_ld = R_out;
    p.cols = R_out;
    p.zv_lo = 1;
    p.zv_hi = 0;
    p.n_stages = mp->n_stages;
    for (int i = 0; i < mp->n_stages; ++i) p.radix[i] = mp->radix[i];
    CK(c, launch_mtd(p, 2, c->stream));
    c->launches++;
    // ---- sum / difference channels, blanking, MATLAB layout ----
    const float* ml = c->s_c.as<float>();
    const float* mr = ml + n_rdm;
    const size_t n_s = (size_t)mtd_fft_num * n_short, n_l = (size_t)mtd_fft_num * fft_num;
    CK(c, c->s_out_re.ensure(std::max<size_t>(n_s + n_l, 1) * sizeof(double)));
    CK(c, c->s_out_im.ensure(std::max<size_t>(n_s + n_l, 1) * sizeof(double)));
    double* d_sum = c->s_out_re.as<double>();
    double* d_diff = c->s_out_im.as<double>();
    if (n_short > 0) {
        CK(c, launch_dmx_combine(ml, mr, R_out, 0, d_sum, d_diff, mtd_fft_num, n_short, n_blank, c->stream));
        c->launches++;
    }
    CK(c, launch_dmx_combine(ml, mr, R_out, n_short, d_sum + n_s, d_diff + n_s, mtd_fft_num, fft_num, n_blank, c->stream));
    c->launches++;
    if (sum_short) CK(c, cudaMemcpyAsync(sum_short, d_sum, n_s * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (diff_short) CK(c, cudaMemcpyAsync(diff_short, d_diff, n_s * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (sum_long) CK(c, cudaMemcpyAsync(sum_long, d_sum + n_s, n_l * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (diff_long) CK(c, cudaMemcpyAsync(diff_long, d_diff + n_s, n_l * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

static int fetch_errflag(rb200_ctx* c, const char* what) {
    CK(c, cudaMemcpyAsync(c->h_counts + 3, c->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->h_counts[3]) return fail(c, RB200_ERR_INDEX, what);
    return RB200_OK;
}

extern "C" int rb200_cfar1d_sub_d(rb200_ctx* c, const double* data, int rows, int cols, int ref, int guard, double T, int method,
                                  double* out) {
    return rb200_cfar1d_fix_d(c, data, rows, cols, ref, guard, T, method, nullptr, 0, nullptr, 0, out);
}

extern "C" int rb200_cfar1d_fix_d(rb200_ctx* c, const double* data, int rows, int cols, int ref, int guard, double T, int method,
                                  const int32_t* rows_fix, int n_rows_fix, const int32_t* cols_fix, int n_cols_fix, double* out) {
    if (!c || !data || !out || rows < 0 || cols < 0 || ref < 1 || guard < 0 || (method != 0 && method != 1) || n_rows_fix < 0 || n_cols_fix < 0)
        return fail(c, RB200_ERR_ARG, "cfar1d: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    const size_t n = (size_t)rows * cols;
    if (n == 0) return RB200_OK;
    for (int i = 0; rows_fix && i < n_rows_fix; ++i)
        if (rows_fix[i] < 1 || rows_fix[i] > rows) return fail(c, RB200_ERR_INDEX, "Function_CFAR1D_sub_fixCells: row index exceeds array bounds");
    for (int i = 0; cols_fix && i < n_cols_fix; ++i)
        if (cols_fix[i] < 1 || cols_fix[i] > cols) return fail(c, RB200_ERR_INDEX, "Function_CFAR1D_sub_fixCells: column index exceeds array bounds");
    CK(c, c->s_in_re.ensure(n * sizeof(double)));
    CK(c, c->s_u8a.ensure(n));
    CK(c, c->s_out_re.ensure(n * sizeof(double)));
    CK(c, cudaMemcpyAsync(c->s_in_re.p, data, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemsetAsync(c->s_u8a.p, 0, n, c->stream));
    CK(c, cudaMemsetAsync(c->errflag.p, 0, sizeof(int), c->stream));
    const int* d_rows = nullptr;
    const int* d_cols = nullptr;
    if (rows_fix || cols_fix) {
        CK(c, c->s_idx.ensure((size_t)(n_rows_fix + n_cols_fix + 1) * sizeof(int)));
        if (rows_fix) {
            CK(c, cudaMemcpyAsync(c->s_idx.p, rows_fix, n_rows_fix * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            d_rows = c->s_idx.as<int>();
        }
        if (cols_fix) {
            CK(c, cudaMemcpyAsync(c->s_idx.as<int>() + n_rows_fix, cols_fix, n_cols_fix * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            d_cols = c->s_idx.as<int>() + n_rows_fix;
        }
    }
    CK(c, launch_cfar1d_f64(c->s_in_re.as<double>(), rows, cols, ref, guard, T, method, d_rows, n_rows_fix, d_cols, n_cols_fix,
                            c->s_u8a.as<uint8_t>(), c->errflag.as<int>(), c->stream));
    c->launches++;
    CK(c, launch_u8_to_d(c->s_u8a.as<uint8_t>(), c->s_out_re.as<double>(), n, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out, c->s_out_re.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return fetch_errflag(c, "Function_CFAR1D_sub: Index exceeds array bounds (axis shorter than 2*(ref+guard))");
}

extern "C" int rb200_execute_cfar_d(rb200_ctx* c, const double* mtd, int V, int R, int ref_r, int guard_r, double t_r, int method_r,
                                    int ref_v, int guard_v, double t_v, int method_v, int n0, int range_stage,
                                    double* out_flag, double* out_flag_v) {
    if (!c || !mtd || !out_flag || V < 1 || R < 1 || ref_r < 1 || ref_v < 1 || guard_r < 0 || guard_v < 0 ||
        (method_r != 0 && method_r != 1) || (method_v != 0 && method_v != 1))
        return fail(c, RB200_ERR_ARG, "execute_cfar: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    // rows n0+2 .. V-n0 (1-based) -> [n0+1, V-n0) 0-based   (CW/executeCFAR.m:23)
    if (n0 < -1 || n0 + 2 < 1 || V - n0 > V) return fail(c, RB200_ERR_INDEX, "executeCFAR: Index in position 1 exceeds array bounds");
    const size_t n = (size_t)V * R;
    CfarParams p;
    memset(&p, 0, sizeof p);
    p.V = V; p.R = R;
    p.v_lo = n0 + 1; p.v_hi = V - n0;
    p.ref_r = ref_r; p.guard_r = guard_r; p.meth_r = method_r;
    p.ref_v = ref_v; p.guard_v = guard_v; p.meth_v = method_v;
    p.range_stage = range_stage ? 1 : 0;
    p.max_det = (int)std::min<size_t>(n, (size_t)1 << 30);
    p.n_lanes = 1; p.cpi0 = 0;
    const int Rw = (R + 31) / 32;
    CK(c, c->s_in_re.ensure(n * sizeof(double)));
    CK(c, c->s_u8a.ensure(n));
    CK(c, c->s_u8b.ensure(n));
    CK(c, c->s_out_re.ensure(n * sizeof(double)));
    CK(c, c->s_a.ensure(n * sizeof(rb200_det)));
    CK(c, c->s_b.ensure((size_t)V * Rw * sizeof(uint32_t)));
    CK(c, cudaMemcpyAsync(c->s_in_re.p, mtd, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemsetAsync(c->s_u8a.p, 0, n, c->stream));
    CK(c, cudaMemsetAsync(c->s_u8b.p, 0, n, c->stream));
    CK(c, cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(int), c->stream));
    CK(c, cudaMemsetAsync(c->errflag.p, 0, sizeof(int), c->stream));
    if (p.v_hi > p.v_lo) {
        CK(c, launch_cfar_f64_colmajor(c->s_in_re.as<double>(), p, t_r, t_v, c->s_a.p, c->counters.as<int>(),
                                       c->s_b.as<uint32_t>(), c->s_u8a.as<uint8_t>(), c->s_u8b.as<uint8_t>(), c->errflag.as<int>(), c->stream));
        c->launches += range_stage ? 2 : 1;
    }
    // rCFARDetect_Flag == 0: cfarResultFlag_Matrix = cfarResultFlag_MatrixV (executeCFAR.m:91)
    const uint8_t* final_flags = range_stage ? c->s_u8a.as<uint8_t>() : c->s_u8b.as<uint8_t>();
    CK(c, launch_u8_to_d(final_flags, c->s_out_re.as<double>(), n, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out_flag, c->s_out_re.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out_flag_v) {
        CK(c, c->s_out_im.ensure(n * sizeof(double)));
        CK(c, launch_u8_to_d(c->s_u8b.as<uint8_t>(), c->s_out_im.as<double>(), n, c->stream));
        c->launches++;
        CK(c, cudaMemcpyAsync(out_flag_v, c->s_out_im.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    return fetch_errflag(c, "executeCFAR: Index exceeds array bounds (CFAR axis shorter than 2*(ref+guard))");
}

// ---------------------------------------------------------------------------------------------
// batched wire-format entry points
// ---------------------------------------------------------------------------------------------
static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int rb200_unpack_ddc_i16(rb200_ctx* c, const int16_t* raw, int n_cpi, float* out_ri) {
    if (!c || !raw || !out_ri || n_cpi < 1) return fail(c, RB200_ERR_ARG, "unpack: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    const rb200_config& k = c->cfg;
    const size_t cells = (size_t)n_cpi * k.n_prt * k.n_range * k.n_lanes;
    CK(c, c->raw.ensure(cells * 4));
    CK(c, c->s_a.ensure(cells * sizeof(float2)));
    CK(c, cudaMemcpyAsync(c->raw.p, raw, cells * 4, cudaMemcpyHostToDevice, c->stream));
    CK(c, launch_unpack(c->raw.as<int16_t>(), c->s_a.as<float2>(), n_cpi * k.n_prt, k.n_prt, k.n_range, k.n_lanes, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out_ri, c->s_a.p, cells * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

// geometry of a DBF-type PRT payload: row width W (bytes per range sample), padded PRT size, complex columns produced
static int dbf24_geometry(rb200_ctx* c, int n, int n_ch, int* W_out, size_t* prt_bytes_out, int* ncol_out) {
    const int osp = 8 - (6 * n_ch) % 8;                                   // FrameDataRead_xzr.m:111
    const int W = 6 * n_ch + osp;
    const size_t sig = (size_t)n * W;                                     // :112
    const size_t prt_bytes = sig + ((sig % 64) ? 64 - sig % 64 : 0);      // :115-119
    // column counts of 1:3:end-3, 2:3:end-2, 3:3:end must agree (MATLAB would raise otherwise)
    const int n1 = (W - 3 >= 1) ? (W - 3 - 1) / 3 + 1 : 0, n2 = (W - 2 >= 2) ? (W - 2 - 2) / 3 + 1 : 0, n3 = (W >= 3) ? (W - 3) / 3 + 1 : 0;
    if (n1 != n2 || n2 != n3 || (n1 % 2)) return fail(c, RB200_ERR_DIM_MISMATCH, "frameDataRead: arrays have incompatible sizes (24-bit word slicing)");
    *W_out = W;
    *prt_bytes_out = prt_bytes;
    *ncol_out = n1 / 2;
    return RB200_OK;
}

extern "C" int rb200_unpack_dbf24(rb200_ctx* c, const uint8_t* bytes, int n_prt, int n, int n_ch, float* out_ri, int* n_columns) {
    if (!c || !bytes || !out_ri || n_prt < 1 || n < 1 || n_ch < 1) return fail(c, RB200_ERR_ARG, "unpack_dbf24: bad argument");
    cudaSetDevice(c->device);
    c->launches = 0;
    int W, ncol;
    size_t prt_bytes;
    int grc = dbf24_geometry(c, n, n_ch, &W, &prt_bytes, &ncol);
    if (grc) return grc;
    if (n_columns) *n_columns = ncol;
    const size_t in_bytes = prt_bytes * n_prt, out_elems = (size_t)ncol * n_prt * n;
    CK(c, c->raw.ensure(in_bytes));
    CK(c, c->s_a.ensure(out_elems * sizeof(float2)));
    CK(c, cudaMemcpyAsync(c->raw.p, bytes, in_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(c, launch_unpack_dbf24(c->raw.as<uint8_t>(), c->s_a.as<float2>(), n_prt, n, ncol, W, prt_bytes, c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out_ri, c->s_a.p, out_elems * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RB200_OK;
}

extern "C" int rb200_motion_para_measure_d(rb200_ctx* c, const double* mtd_sum, const double* mtd_diff, const double* flags, int V, int R,
                                           int extra, const double* r_scale, double delta_r, int r_times, const double* v_scale,
                                           double delta_v, int v_times, const double* k_values, int k_rows, int k_cols,
                                           double beam_pos, double beam_step, int fre_ind, double ele_comp, double ele_err, int n0,
                                           double* out_r, double* out_v, double* out_e, int capacity, int* n_out) {
    if (!c || !mtd_sum || !mtd_diff || !flags || !r_scale || !v_scale || !k_values || !n_out || V < 1 || R < 1 || extra < 0 ||
        r_times < 1 || v_times < 1 || capacity < 0 || k_rows < 1 || k_cols < 1)
        return fail(c, RB200_ERR_ARG, "motionParaMeasure: bad argument");
    if (extra > 16) return fail(c, RB200_ERR_UNSUPPORTED, "motionParaMeasure: extraDots > 16 is not supported");
    if (fre_ind < 0 || fre_ind >= k_rows || beam_pos < 0 || (int)beam_pos >= k_cols)
        return fail(c, RB200_ERR_INDEX, "motionParaMeasure: kValues(freInd+1, beamPosNum+1) exceeds array bounds");
    if (2 * extra + 1 > R || 2 * extra + 1 > V - 2 * n0 - 1)
        return fail(c, RB200_ERR_INDEX, "motionParaMeasure: Index exceeds array bounds (fewer cells than 2*extraDots+1)");
    cudaSetDevice(c->device);
    c->launches = 0;
    const size_t n = (size_t)V * R;
    // device copies: [sum | diff | flags | rScale | vScale | kValues]
    const size_t nd = 3 * n + R + V + (size_t)k_rows * k_cols;
    CK(c, c->s_in_re.ensure(nd * sizeof(double)));
    double* d = c->s_in_re.as<double>();
    double *d_sum = d, *d_diff = d + n, *d_flags = d + 2 * n, *d_rs = d + 3 * n, *d_vs = d_rs + R, *d_k = d_vs + V;
    CK(c, cudaMemcpyAsync(d_sum, mtd_sum, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(d_diff, mtd_diff, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(d_flags, flags, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(d_rs, r_scale, R * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(d_vs, v_scale, V * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpyAsync(d_k, k_values, (size_t)k_rows * k_cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(c, c->s_idx.ensure((size_t)(R + 1) * sizeof(int)));
    int* col_start = c->s_idx.as<int>();
    CK(c, cudaMemsetAsync(c->errflag.p, 0, sizeof(int), c->stream));
    CK(c, launch_flag_compaction(d_flags, V, R, col_start, col_start + R, c->stream));
    c->launches += 2;
    CK(c, cudaMemcpyAsync(c->h_counts, col_start + R, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    const int total = c->h_counts[0];
    *n_out = total;
    if (total == 0) return RB200_OK;
    if (total > capacity) return fail(c, RB200_ERR_OVERFLOW, "motionParaMeasure: more flagged cells than the output capacity");
    if (!out_r || !out_v || !out_e) return fail(c, RB200_ERR_ARG, "motionParaMeasure: output arrays are NULL");
    CK(c, c->s_out_re.ensure((size_t)3 * total * sizeof(double)));
    double* o = c->s_out_re.as<double>();
    CK(c, launch_measure(d_sum, d_diff, d_flags, d_rs, d_vs, d_k, k_rows, V, R, extra, r_times, v_times, n0, delta_r, delta_v, beam_pos,
                         beam_step, fre_ind, ele_comp, ele_err, col_start, o, o + total, o + 2 * total, c->errflag.as<int>(), c->stream));
    c->launches++;
    CK(c, cudaMemcpyAsync(out_r, o, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(out_v, o + total, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(out_e, o + 2 * total, total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return fetch_errflag(c, "motionParaMeasure: Index exceeds array bounds");
}

static int chunk_size(const rb200_ctx* c, bool host_staged = false) {
    int g = c->env.chunk ? c->env.chunk : c->cfg.chunk_cpi;
    if (g <= 0) {
        // default for device-resident batches: ~2.1 GB of raw+PC+RDM per chunk (S3: 32 CPIs).  Every launch of the persistent
        // kernels costs a prologue and a partly filled last round (measured with pcw_kernel, S3, stage times per CPI:
        // chunk 8 -> 14.5 / 9.7 / 2.3 us, chunk 32 -> 13.0 / 8.8 / 1.4 us), and the PC intermediate is not L2-resident
        // at any practical chunk size (profiles/README.md)
        const double per_cpi = (double)c->cfg.n_prt * c->cfg.n_range * (c->dbf_beams ? c->dbf_beams : c->cfg.n_lanes) * 16.0;
        g = (int)std::floor(2150e6 / per_cpi);
        // host buffers: the call is PCIe-bound and the first H2D / last D2H of a call cannot overlap anything, so keep
        // the chunks small (S3: 4 CPIs; measured: 2.63 k -> 2.71 k CPI/s end to end against 8-CPI chunks)
        if (host_staged) g = std::max(1, g / 8);
    }
    // grid.y carries (CPIs x lanes) slabs and, on the DBF path, (CPIs x PRTs) groups: keep both inside the 65535 limit
    const int lanes = std::max(1, std::max(c->cfg.n_lanes, c->dbf_beams));
    g = std::min(g, 65535 / lanes);
    if (c->dbf_beams) g = std::min(g, 65535 / std::max(1, c->cfg.n_prt));
    return std::max(1, std::min(g, c->cfg.max_cpi));
}

// raw_host / rdm_host (optional): host buffers staged chunk by chunk on the chunk's own stream, so the H2D copy
// of chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i (three slots, two copy engines).
// dbf24_ch > 0: the input is the DBF-type 24-bit payload of `dbf24_ch` channels per PRT (row f1, FrameDataRead_xzr.m:111-119,
// 130-135,163) whose complex columns are the cfg.n_lanes lanes of the chain; raw_* then point at bytes.
static int chain_enqueue(rb200_ctx* c, const int16_t* raw_dev, int n_cpi, float* rdm_dev, cudaStream_t st,
                         const int16_t* raw_host = nullptr, float* rdm_host = nullptr, int dbf24_ch = 0) {
    const rb200_config& k = c->cfg;
    const int P = k.n_prt, R = k.n_range;
    const int Cin = k.n_lanes;                                   // channels interleaved in the wire format
    const int C = c->dbf_beams ? c->dbf_beams : Cin;             // lanes downstream of the (optional) beam former
    int d24_W = 0, d24_ncol = 0;
    size_t d24_prt_bytes = 0;
    if (dbf24_ch > 0) {
        if (c->dbf_beams) return fail(c, RB200_ERR_ARG, "chain_dbf24: DBF-type data is already beam-formed (clear rb200_set_dbf first)");
        int grc = dbf24_geometry(c, R, dbf24_ch, &d24_W, &d24_prt_bytes, &d24_ncol);
        if (grc) return grc;
        if (d24_ncol != C) return fail(c, RB200_ERR_DIM_MISMATCH, "chain_dbf24: the payload's complex column count differs from n_lanes");
    }
    if (n_cpi < 1 || n_cpi > k.max_cpi) return fail(c, RB200_ERR_ARG, "chain: n_cpi must be in 1..max_cpi");
    if (k.cfar_n0 < 0) return fail(c, RB200_ERR_INDEX, "executeCFAR: Index in position 1 exceeds array bounds (MTD_0_num < 0)");
    int G = chunk_size(c, raw_host != nullptr || rdm_host != nullptr);
    // a batch that fits one default chunk still runs as two, so that the tail of one chunk's kernels overlaps the head of the
    // next on the slot streams (measured: 64 CPIs as 2 x 32 run 3 % faster than as 1 x 64)
    if (!c->env.chunk && c->cfg.chunk_cpi <= 0 && n_cpi >= 16 && n_cpi < 2 * G) G = (n_cpi + 1) / 2;
    const size_t cpi_cells = (size_t)P * R * C;
    const size_t raw_cells = (size_t)P * R * Cin;
    const size_t raw_cpi_bytes = dbf24_ch > 0 ? (size_t)P * d24_prt_bytes : raw_cells * 4;   // input bytes of one CPI
    const bool planar_in = c->dbf_beams > 0 || dbf24_ch > 0;    // a first kernel produces planar [cpi][lane][prt][range] lines
    const int Rw = (R + 31) / 32;
    CK(c, c->dets_v.ensure((size_t)k.max_det * sizeof(rb200_det)));
    CK(c, c->dets_2d.ensure((size_t)k.max_det * sizeof(rb200_det)));
    float* rdm_base = rdm_dev;
    CfarParams cp;
    memset(&cp, 0, sizeof cp);
    cp.V = P; cp.R = R;
    cp.v_lo = k.cfar_n0 + 1; cp.v_hi = P - k.cfar_n0;
    cp.ref_r = k.cfar_ref_r; cp.guard_r = k.cfar_guard_r; cp.meth_r = k.cfar_method_r;
    cp.ref_v = k.cfar_ref_v; cp.guard_v = k.cfar_guard_v; cp.meth_v = k.cfar_method_v;
    cp.range_stage = k.cfar_range_stage ? 1 : 0;
    cp.max_det = k.max_det;
    cp.n_lanes = C;
    cp.segs = c->cfar_segs;
    for (int i = 0; i < cp.segs.n; ++i)
        if (cp.segs.hi[i] > R) return fail(c, RB200_ERR_INDEX, "fun_CFARflag: Index in position 2 exceeds array bounds (range segment past the PRT length)");
    c->launches = 0;
    CK(c, cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(int), st));
    CK(c, cudaMemsetAsync(c->errflag.p, 0, sizeof(int), st));
    const bool fused = !c->env.no_fused && mtd64_fused_supported(P, k.cfar_ref_v, k.cfar_guard_v, k.cfar_n0, k.mti_lag);
    Mtd64Params m64;
    memset(&m64, 0, sizeof m64);
    int n_slots = 1;
    c->last_was_mega = false;
    c->last_was_onepass = false;
    // ---- single-pass kernel (onepass_kernel.cu): P = 64, 16 interleaved int16 lanes, one matched-filter segment covering the
    //      whole PRT, default velocity windows; the pulse-compressed intermediate stays in shared memory
    bool onepass = false;
    int op_V = 0, op_tiles = 0;
    if (fused && c->env.onepass && !c->keep_pc && C == 16 && Cin == 16 && !planar_in && c->gain_n == 0 && c->plan.valid &&
        c->plan.segs.size() == 1 && c->plan.classes.size() == 1 && c->plan.classes[0].nt == 256 && (R % 4) == 0) {
        const PcSegDev& d = c->plan.segs[0].d;
        const bool whole = d.in_start == 0 && d.out_start == 0 && d.in_len == R && d.out_len == R && d.pre == 0 && d.rot == 0 && d.h_off == 0;
        const int V = onepass_tile_valid(d.n_taps);
        const bool aligned = (raw_host || (reinterpret_cast<uintptr_t>(raw_dev) & 15) == 0) && (!rdm_dev || (reinterpret_cast<uintptr_t>(rdm_dev) & 15) == 0);
        if (whole && V >= 64 && d.V >= V && aligned) {
            onepass = true;
            op_V = V;
            op_tiles = (R + V - 1) / V;
        }
    }
    // ---- fused persistent kernel for the whole batch (chain64_kernel.cu): device-resident input and output, 16 channels,
    //      one 256-sample tile class covering the whole PRT
    if (fused && raw_dev && rdm_dev && !raw_host && !rdm_host && C == 16 && !planar_in && c->plan.valid && c->plan.classes.size() == 1 &&
        c->plan.classes[0].nt == 256 && c->plan.max_in_end <= R && c->plan.max_out_end <= R && (reinterpret_cast<uintptr_t>(raw_dev) & 15) == 0 &&
        c->env.mega) {      // opt-in: measured slower than the slot pipeline on B200 (profiles/README.md)
        bool direct = false, covered = true;
        for (auto& sp : c->plan.segs) direct |= (sp.nt == 0 && sp.d.out_len > 0);
        int cur = 0;
        for (auto& rg : c->plan.out_ranges) { if (rg.first > cur) covered = false; cur = std::max(cur, rg.second); }
        if (cur < R) covered = false;
        if (!direct && covered) {
            MtdPlan* mp = nullptr;
            int rc = get_mtd_plan(c, P, k.kaiser_beta, &mp);
            if (rc) return rc;
            int zlo, zhi;
            if (zero_v_rows(P, k.zero_v_div, &zlo, &zhi)) return fail(c, RB200_ERR_INDEX, "fun_0v_pressing: Index in position 1 is invalid");
            const int ring_slots = 3;
            rb200_ctx::Slot& sl = c->slots[0];
            CK(c, c->ring.ensure((size_t)ring_slots * cpi_cells * sizeof(float2)));
            CK(c, c->colmask.ensure((size_t)n_cpi * C * R * sizeof(unsigned long long)));
            CK(c, sl.vlist.ensure((size_t)k.max_det * sizeof(rb200_det)));
            CK(c, c->megactr.ensure((size_t)(1 + 2 * n_cpi) * sizeof(int)));
            Chain64Params q;
            memset(&q, 0, sizeof q);
            q.pc.in = raw_dev;
            q.pc.hperm = c->plan.hperm.as<float2>();
            q.pc.tw = c->plan.classes[0].tw.as<float2>();
            q.pc.tiles = c->plan.classes[0].tiles.as<int2>();
            q.pc.gain = c->gain_n ? c->gain.as<float>() : nullptr;
            for (size_t i = 0; i < c->plan.segs.size(); ++i) q.pc.segs[i] = c->plan.segs[i].d;
            q.pc.R = R; q.pc.R_out = R; q.pc.C = C; q.pc.P = P;
            for (int i = 0; i < 64; ++i) {
                q.m.win[i] = mp->h_window[i];
                q.m.keep[i] = (i >= zlo && i <= zhi) ? 0.f : 1.f;
            }
            q.m.out = rdm_dev;
            q.m.in_ld = q.m.out_ld = q.m.cols = R;
            q.m.meth_v = k.cfar_method_v;
            q.m.tv_over_ref = (float)(k.cfar_t_v / k.cfar_ref_v);
            q.m.dets = sl.vlist.p;
            q.m.det_count = sl.count.as<int>();
            q.m.colmask = c->colmask.as<unsigned long long>();
            q.m.cols_ld = R;
            q.m.max_det = k.max_det;
            q.m.n_lanes = C;
            q.m.segs = c->cfar_segs;
            q.m.cpi0 = 0;
            q.n_cpi = n_cpi;
            q.n_tiles = c->plan.classes[0].n_tiles;
            q.n_pc_items = P * q.n_tiles;
            q.n_mtd_items = C * ((R + 255) / 256);
            q.ring = c->ring.as<float2>();
            q.ring_slots = ring_slots;
            q.ring_stride = cpi_cells;
            q.work_counter = c->megactr.as<int>();
            q.pc_done = c->megactr.as<int>() + 1;
            q.mtd_done = c->megactr.as<int>() + 1 + n_cpi;
            q.err_flag = c->errflag.as<int>();
            CK(c, cudaMemsetAsync(c->megactr.p, 0, (size_t)(1 + 2 * n_cpi) * sizeof(int), st));
            CK(c, cudaEventRecord(c->ev0, st));
            const bool timed = c->stage_timing && c->stage_used + 4 <= 65536;
            if (timed) { stage_event(c, st); c->stage_cpis.push_back(n_cpi); }
            CK(c, launch_chain64(q, c->n_sms, st));
            c->launches++;
            if (timed) { stage_event(c, st); stage_event(c, st); }
            cp.cpi0 = 0;
            CK(c, launch_cfar_r64(rdm_dev, cp, (float)k.cfar_t_r, sl.vlist.p, sl.count.as<int>(), c->dets_v.p, c->dets_2d.p,
                                  c->counters.as<int>(), c->colmask.as<unsigned long long>(), R, c->errflag.as<int>(), c->n_sms * 2, st));
            c->launches++;
            if (timed) stage_event(c, st);
            CK(c, cudaEventRecord(c->ev1, st));
            c->have_timing = true;
            c->last_chunk_cpis = 0;
            c->last_pc = nullptr;
            c->last_was_mega = true;
            return RB200_OK;
        }
    }
    if (fused) {
        MtdPlan* mp = nullptr;
        int rc = get_mtd_plan(c, P, k.kaiser_beta, &mp);
        if (rc) return rc;
        int zlo, zhi;
        if (zero_v_rows(P, k.zero_v_div, &zlo, &zhi)) return fail(c, RB200_ERR_INDEX, "fun_0v_pressing: Index in position 1 is invalid");
        for (int i = 0; i < 64; ++i) {
            m64.win[i] = mp->h_window[i];
            m64.keep[i] = (i >= zlo && i <= zhi) ? 0.f : 1.f;
        }
        m64.in_ld = m64.out_ld = m64.cols = R;
        m64.meth_v = k.cfar_method_v;
        m64.tv_over_ref = (float)(k.cfar_t_v / k.cfar_ref_v);
        m64.cols_ld = R;
        m64.max_det = k.max_det;
        m64.n_lanes = C;
        m64.segs = c->cfar_segs;
        n_slots = c->env.slots ? c->env.slots : 3;
        n_slots = std::max(1, std::min(n_slots, (int)rb200_ctx::kMaxSlots));
        if (c->stage_timing) n_slots = 1;                      // per-stage events need the chunks serialised
        n_slots = std::min(n_slots, (n_cpi + G - 1) / G);
        for (int i = 0; i < n_slots; ++i) {
            rb200_ctx::Slot& sl = c->slots[i];
            if (!onepass) CK(c, sl.pc.ensure((size_t)G * cpi_cells * sizeof(float2)));
            else {
                const size_t pb = onepass_planar_bytes(G, R, op_V);
                CK(c, sl.op_planar.ensure(pb));
                if (sl.op_planar_cap != sl.op_planar.cap || sl.op_R != R || sl.op_V != op_V) {
                    CK(c, onepass_planar_init(sl.op_planar.p, sl.op_planar.cap, st));
                    sl.op_planar_cap = sl.op_planar.cap;
                    sl.op_R = R;
                    sl.op_V = op_V;
                }
            }
            CK(c, sl.colmask.ensure((size_t)G * C * R * sizeof(unsigned long long)));
            CK(c, sl.vlist.ensure((size_t)k.max_det * sizeof(rb200_det)));
        }
    } else {
        CK(c, c->pc.ensure((size_t)G * cpi_cells * sizeof(float2)));
        CK(c, c->vmask.ensure((size_t)G * C * P * Rw * sizeof(uint32_t)));
    }
    for (int i = 0; i < n_slots; ++i) {
        if (raw_host) CK(c, c->slots[i].raw.ensure((size_t)G * raw_cpi_bytes));
        if (planar_in) CK(c, c->slots[i].beams.ensure((size_t)G * cpi_cells * sizeof(float2)));
        // host output, or no output at all: every slot owns the RDM of its chunk (the range stage and the detection
        // amplitudes of chunk i read it while chunk i+1 is already being transformed on another slot stream)
        if (rdm_host || !rdm_dev) CK(c, c->slots[i].rdm.ensure((size_t)G * cpi_cells * sizeof(float)));
    }
    // RB200_SPLIT=n1 (experiment): pcw_kernel of every chunk on n1 SMs (stream A), mtd64_tma + range stage on the other SMs
    // (stream B) at the same time; K2 fetches a tile only when K1 has counted the tile's CPI complete, so it reads the
    // intermediate out of L2.  Device-resident input and output, default K1 / K2 kernels only.
    const int n_chunks_total = (n_cpi + G - 1) / G;
    const bool split_on = fused && !onepass && c->env.split > 0 && c->env.split < c->n_sms && raw_dev && !raw_host && !rdm_host && rdm_dev &&
                          !c->stage_timing && dbf24_ch == 0 && !c->dbf_beams && !planar_in && (R % 2) == 0 && !c->env.no_tma_mtd &&
                          pcw_eligible(c, c->plan, true, raw_dev, R, C);
    int* split_done = nullptr;
    int* split_started = nullptr;
    if (split_on) {
        CK(c, c->split_ctr.ensure((size_t)(n_cpi + n_chunks_total) * sizeof(int)));
        CK(c, cudaMemsetAsync(c->split_ctr.p, 0, (size_t)(n_cpi + n_chunks_total) * sizeof(int), st));
        split_done = c->split_ctr.as<int>();
        split_started = split_done + n_cpi;
    }
    const int n_fork = split_on ? 2 : n_slots;
    CK(c, cudaEventRecord(c->ev0, st));
    if (n_fork > 1) {
        CK(c, cudaEventRecord(c->fork_ev, st));
        for (int i = 0; i < n_fork; ++i) CK(c, cudaStreamWaitEvent(c->slots[i].stream, c->fork_ev, 0));
    }
    int chunk_idx = 0;
    for (int c0 = 0; c0 < n_cpi; c0 += G, ++chunk_idx) {
        const int g = std::min(G, n_cpi - c0);
        rb200_ctx::Slot& sl = c->slots[chunk_idx % n_slots];
        cudaStream_t cs = (fused && n_slots > 1) ? sl.stream : st;
        const int16_t* raw_chunk =
            raw_dev ? reinterpret_cast<const int16_t*>(reinterpret_cast<const uint8_t*>(raw_dev) + (size_t)c0 * raw_cpi_bytes) : nullptr;
        if (raw_host) {
            CK(c, cudaMemcpyAsync(sl.raw.p, reinterpret_cast<const uint8_t*>(raw_host) + (size_t)c0 * raw_cpi_bytes, (size_t)g * raw_cpi_bytes,
                                  cudaMemcpyHostToDevice, cs));
            raw_chunk = sl.raw.as<int16_t>();
        }
        float* rdm_chunk = rdm_dev ? rdm_base + (size_t)c0 * cpi_cells : sl.rdm.as<float>();
        float2* pc_buf = fused ? sl.pc.as<float2>() : c->pc.as<float2>();
        const bool timed = c->stage_timing && c->stage_used + 4 <= 65536;
        if (timed) { stage_event(c, cs); c->stage_cpis.push_back(g); }
        int rc;
        if (onepass) {
            OnePassParams op;
            memset(&op, 0, sizeof op);
            op.rdm = rdm_chunk;
            op.hperm = c->plan.hperm.as<float2>();
            op.tw = c->plan.classes[0].tw.as<float2>();
            op.colmask = sl.colmask.as<unsigned long long>();
            op.dets = sl.vlist.p;
            op.det_count = sl.count.as<int>();
            op.err_flag = c->errflag.as<int>();
            op.max_det = k.max_det;
            op.R = R; op.V = op_V; op.n_tiles = op_tiles; op.n_cpi = g;
            op.cpi0 = c0;
            op.meth_v = k.cfar_method_v;
            op.tv_over_ref = m64.tv_over_ref;
            for (int i = 0; i < 64; ++i) {
                op.win[i] = m64.win[i];
                op.keep[i] = m64.keep[i];
            }
            op.segs = c->cfar_segs;
            op.dbg = c->env.op_dbg;
            CK(c, launch_deinterleave(raw_chunk, sl.op_planar.p, g, R, op_V, cs));
            c->launches++;
            CK(c, launch_onepass(op, sl.op_planar.p, c->n_sms, cs));
            c->launches++;
            if (timed) { stage_event(c, cs); stage_event(c, cs); }      // "pc" = the whole single-pass kernel, "mtd" = 0
            cp.cpi0 = c0;
            CK(c, launch_cfar_r64(rdm_chunk, cp, (float)k.cfar_t_r, sl.vlist.p, sl.count.as<int>(), c->dets_v.p, c->dets_2d.p,
                                  c->counters.as<int>(), sl.colmask.as<unsigned long long>(), R, c->errflag.as<int>(), c->n_sms * 16, cs));
            c->launches++;
            if (rdm_host)
                CK(c, cudaMemcpyAsync(rdm_host + (size_t)c0 * cpi_cells, rdm_chunk, (size_t)g * cpi_cells * sizeof(float), cudaMemcpyDeviceToHost, cs));
            if (timed) stage_event(c, cs);
            c->last_chunk_cpis = g;
            c->last_pc = nullptr;
            c->last_was_onepass = true;
            continue;
        }
        if (dbf24_ch > 0) {
            // f1: 24-bit DBF-type payload -> planar [cpi][beam][prt][range], one launch per CPI of the chunk
            for (int q = 0; q < g; ++q) {
                CK(c, launch_unpack_dbf24(reinterpret_cast<const uint8_t*>(raw_chunk) + (size_t)q * raw_cpi_bytes,
                                          sl.beams.as<float2>() + (size_t)q * cpi_cells, P, R, C, d24_W, d24_prt_bytes, cs));
                c->launches++;
            }
            rc = run_pc(c, c->plan, false, sl.beams.p, pc_buf, R, R, 1, P, 0, g * C * P, c->gain_n ? c->gain.as<float>() : nullptr, cs);
        } else if (c->dbf_beams) {
            // f1: beams = sig_C * W.' fused with the unpack, then planar pulse compression over cpi x beam x PRT lines
            CK(c, launch_dbf(raw_chunk, sl.beams.as<float2>(), c->dbf_w.as<float2>(), C, Cin, g * P, P, R, cs));
            c->launches++;
            rc = run_pc(c, c->plan, false, sl.beams.p, pc_buf, R, R, 1, P, 0, g * C * P, c->gain_n ? c->gain.as<float>() : nullptr, cs);
        } else if (split_on) {
            cudaStream_t sa = c->slots[0].stream, sb = c->slots[1].stream;
            if (chunk_idx >= n_slots) CK(c, cudaStreamWaitEvent(sa, sl.done, 0));        // the slot's buffers are free again
            rc = run_pc(c, c->plan, true, raw_chunk, pc_buf, R, R, C, P, g * P, 0, c->gain_n ? c->gain.as<float>() : nullptr, sa,
                        split_done + c0, split_started + chunk_idx, c->env.split);
            if (rc) return rc;
            cp.cpi0 = c0;
            m64.in = pc_buf;
            m64.out = rdm_chunk;
            m64.cpi0 = c0;
            m64.dets = sl.vlist.p;
            m64.det_count = sl.count.as<int>();
            m64.colmask = sl.colmask.as<unsigned long long>();
            m64.wait_done = split_done + c0;
            m64.wait_target = c->plan.classes[0].n_tiles * P * 4;          // four warp tasks per (PRT, tile) item
            m64.err_flag = c->errflag.as<int>();
            CK(c, launch_wait_flag(split_started + chunk_idx, c->errflag.as<int>(), sb));
            CK(c, launch_mtd64_tma(m64, g * C, c->n_sms - c->env.split, c->mtd_ctas_per_sm, sb));
            CK(c, launch_cfar_r64(rdm_chunk, cp, (float)k.cfar_t_r, sl.vlist.p, sl.count.as<int>(), c->dets_v.p, c->dets_2d.p,
                                  c->counters.as<int>(), sl.colmask.as<unsigned long long>(), R, c->errflag.as<int>(), c->n_sms * 16, sb));
            CK(c, cudaEventRecord(sl.done, sb));
            c->launches += 4;
            c->last_chunk_cpis = g;
            c->last_pc = pc_buf;
            continue;
        } else {
            rc = run_pc(c, c->plan, true, raw_chunk, pc_buf, R, R, C, P, g * P, 0, c->gain_n ? c->gain.as<float>() : nullptr, cs);
        }
        if (rc) return rc;
        if (timed) stage_event(c, cs);
        cp.cpi0 = c0;
        if (fused) {
            // K2 + velocity CFAR in one kernel (register-resident Doppler columns), then the sparse range stage
            m64.in = pc_buf;
            m64.out = rdm_chunk;
            m64.cpi0 = c0;
            m64.dets = sl.vlist.p;
            m64.det_count = sl.count.as<int>();
            m64.colmask = sl.colmask.as<unsigned long long>();
            if ((R % 2) == 0 && !c->env.no_tma_mtd) CK(c, launch_mtd64_tma(m64, g * C, c->n_sms, c->mtd_ctas_per_sm, cs));
            else CK(c, launch_mtd64(m64, g * C, true, cs));
            c->launches++;
            if (timed) stage_event(c, cs);
            CK(c, launch_cfar_r64(rdm_chunk, cp, (float)k.cfar_t_r, sl.vlist.p, sl.count.as<int>(), c->dets_v.p, c->dets_2d.p,
                                  c->counters.as<int>(), sl.colmask.as<unsigned long long>(), R, c->errflag.as<int>(), c->n_sms * 16, cs));
            c->launches++;
        } else {
            // hits of this chunk start where the list currently ends
            if (cp.v_hi > cp.v_lo)
                CK(c, cudaMemcpyAsync(c->counters.as<int>() + 2, c->counters.as<int>() + 0, sizeof(int), cudaMemcpyDeviceToDevice, cs));
            const bool fuse_v = cp.v_hi > cp.v_lo && mtd_fast_fuses_cfar(P) && (cp.v_hi - cp.v_lo) >= 2 * (cp.ref_v + cp.guard_v) &&
                                !c->env.no_fused_v && !(c->env.mtd_tc && P == 64);
            FusedV fvp = {&cp, (float)k.cfar_t_v, c->dets_v.p, c->counters.as<int>() + 0, c->vmask.as<uint32_t>(), c->errflag.as<int>()};
            rc = run_mtd(c, pc_buf, rdm_chunk, P, R, R, R, g * C, k.kaiser_beta, k.zero_v_div, k.mti_lag, cs, fuse_v ? &fvp : nullptr);
            if (rc) return rc;
            if (timed) stage_event(c, cs);
            if (cp.v_hi > cp.v_lo) {
                if (fuse_v) {
                    CK(c, launch_cfar_r_f32(rdm_chunk, cp, (float)k.cfar_t_r, c->dets_v.p, c->counters.as<int>() + 0, c->dets_2d.p,
                                            c->counters.as<int>() + 1, c->vmask.as<uint32_t>(), c->errflag.as<int>(), cs));
                    c->launches += cp.range_stage ? 1 : 0;
                } else {
                    CK(c, launch_cfar_f32(rdm_chunk, cp, (float)k.cfar_t_r, (float)k.cfar_t_v, g * C, c->dets_v.p, c->counters.as<int>() + 0,
                                          c->dets_2d.p, c->counters.as<int>() + 1, c->vmask.as<uint32_t>(), nullptr, nullptr,
                                          c->errflag.as<int>(), cs));
                    c->launches += cp.range_stage ? 2 : 1;
                }
            }
        }
        if (rdm_host)
            CK(c, cudaMemcpyAsync(rdm_host + (size_t)c0 * cpi_cells, rdm_chunk, (size_t)g * cpi_cells * sizeof(float), cudaMemcpyDeviceToHost, cs));
        if (timed) stage_event(c, cs);
        c->last_chunk_cpis = g;
        c->last_pc = pc_buf;
    }
    if (n_fork > 1) {
        for (int i = 0; i < n_fork; ++i) {
            CK(c, cudaEventRecord(c->slots[i].done, c->slots[i].stream));
            CK(c, cudaStreamWaitEvent(st, c->slots[i].done, 0));
        }
    }
    CK(c, cudaEventRecord(c->ev1, st));
    c->have_timing = true;
    return RB200_OK;
}

extern "C" int rb200_chain_enqueue(rb200_ctx* c, const int16_t* raw_dev, int n_cpi, float* rdm_dev, void* stream) {
    if (!c || !raw_dev) return fail(c, RB200_ERR_ARG, "chain_enqueue: bad argument");
    cudaSetDevice(c->device);
    c->last_stream = stream ? (cudaStream_t)stream : c->stream;
    return chain_enqueue(c, raw_dev, n_cpi, rdm_dev, c->last_stream);
}

static int chain_fetch(rb200_ctx* c, rb200_det* dets, bool dets_on_device, int* n_det, cudaStream_t st) {
    const int cap = c->cfg.max_det;
    CK(c, cudaMemcpyAsync(c->h_counts, c->counters.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_counts + 3, c->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    const int nv = c->h_counts[0], n2 = c->cfg.cfar_range_stage ? c->h_counts[1] : 0;
    if (n_det) *n_det = nv + n2;
    if (c->h_counts[3] == 2) return fail(c, RB200_ERR_CUDA, "fused chain kernel: a dependency wait timed out (internal error)");
    if (c->h_counts[3]) return fail(c, RB200_ERR_INDEX, "executeCFAR: Index exceeds array bounds (CFAR axis shorter than 2*(ref+guard))");
    if (dets) {
        // 2-D records first (the product), then velocity-stage records; at most max_det in total
        const int k2 = std::min(n2, cap);
        const int kv = std::min(std::min(nv, cap), cap - k2);
        const cudaMemcpyKind kind = dets_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        if (k2) CK(c, cudaMemcpyAsync(dets, c->dets_2d.p, (size_t)k2 * sizeof(rb200_det), kind, st));
        if (kv) CK(c, cudaMemcpyAsync(dets + k2, c->dets_v.p, (size_t)kv * sizeof(rb200_det), kind, st));
        CK(c, cudaStreamSynchronize(st));
    }
    if (nv > cap || nv + n2 > cap) return fail(c, RB200_ERR_OVERFLOW, "detection list truncated: raise rb200_config.max_det");
    return RB200_OK;
}

extern "C" int rb200_chain_fetch(rb200_ctx* c, rb200_det* dets_host, int* n_det) {
    if (!c) return RB200_ERR_ARG;
    cudaSetDevice(c->device);
    // same stream as the enqueue: the copies are ordered after the chain's kernels
    return chain_fetch(c, dets_host, dets_host && is_device_ptr(dets_host), n_det, c->last_stream ? c->last_stream : c->stream);
}

extern "C" int rb200_chain_dets_device(rb200_ctx* c, const rb200_det** dets_2d, int* n_2d, const rb200_det** dets_v, int* n_v) {
    if (!c) return RB200_ERR_ARG;
    cudaSetDevice(c->device);
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    CK(c, cudaMemcpyAsync(c->h_counts, c->counters.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(c, cudaMemcpyAsync(c->h_counts + 3, c->errflag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    const int cap = c->cfg.max_det;
    const int nv = c->h_counts[0], n2 = c->cfg.cfar_range_stage ? c->h_counts[1] : 0;
    if (dets_2d) *dets_2d = c->dets_2d.as<rb200_det>();
    if (dets_v) *dets_v = c->dets_v.as<rb200_det>();
    if (n_2d) *n_2d = std::min(n2, cap);
    if (n_v) *n_v = std::min(nv, cap);
    if (c->h_counts[3] == 2) return fail(c, RB200_ERR_CUDA, "fused chain kernel: a dependency wait timed out (internal error)");
    if (c->h_counts[3]) return fail(c, RB200_ERR_INDEX, "executeCFAR: Index exceeds array bounds (CFAR axis shorter than 2*(ref+guard))");
    if (nv > cap || n2 > cap) return fail(c, RB200_ERR_OVERFLOW, "detection list truncated: raise rb200_config.max_det");
    return RB200_OK;
}

extern "C" int rb200_chain_i16(rb200_ctx* c, const int16_t* raw, int n_cpi, float* rdm_out, rb200_det* dets, int* n_det, void* stream) {
    if (!c || !raw) return fail(c, RB200_ERR_ARG, "chain: bad argument");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const rb200_config& k = c->cfg;
    if (n_cpi < 1 || n_cpi > k.max_cpi) return fail(c, RB200_ERR_ARG, "chain: n_cpi must be in 1..max_cpi");
    const bool raw_on_dev = is_device_ptr(raw);
    const bool rdm_on_dev = rdm_out && is_device_ptr(rdm_out);
    int rc = chain_enqueue(c, raw_on_dev ? raw : nullptr, n_cpi, rdm_on_dev ? rdm_out : nullptr, st,
                           raw_on_dev ? nullptr : raw, (rdm_out && !rdm_on_dev) ? rdm_out : nullptr);
    if (rc) return rc;
    return chain_fetch(c, dets, dets && is_device_ptr(dets), n_det, st);
}

extern "C" int rb200_chain_dbf24(rb200_ctx* c, const uint8_t* payload, int n_ch, int n_cpi, float* rdm_out, rb200_det* dets, int* n_det,
                                 void* stream) {
    if (!c || !payload || n_ch < 1) return fail(c, RB200_ERR_ARG, "chain_dbf24: bad argument");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const rb200_config& k = c->cfg;
    if (n_cpi < 1 || n_cpi > k.max_cpi) return fail(c, RB200_ERR_ARG, "chain: n_cpi must be in 1..max_cpi");
    const bool raw_on_dev = is_device_ptr(payload);
    const bool rdm_on_dev = rdm_out && is_device_ptr(rdm_out);
    const int16_t* raw = reinterpret_cast<const int16_t*>(payload);
    int rc = chain_enqueue(c, raw_on_dev ? raw : nullptr, n_cpi, rdm_on_dev ? rdm_out : nullptr, st, raw_on_dev ? nullptr : raw,
                           (rdm_out && !rdm_on_dev) ? rdm_out : nullptr, n_ch);
    if (rc) return rc;
    return chain_fetch(c, dets, dets && is_device_ptr(dets), n_det, st);
}

extern "C" int rb200_set_debug_keep_pc(rb200_ctx* c, int on) {
    if (!c) return RB200_ERR_ARG;
    c->keep_pc = on != 0;
    return RB200_OK;
}

extern "C" int rb200_debug_fetch_pc(rb200_ctx* c, int cpi_in_chunk, float* out_ri) {
    if (!c || !out_ri || cpi_in_chunk < 0 || cpi_in_chunk >= c->last_chunk_cpis) return fail(c, RB200_ERR_ARG, "debug_fetch_pc: bad argument");
    cudaSetDevice(c->device);
    const size_t cpi_cells = (size_t)c->cfg.n_prt * c->cfg.n_range * (c->dbf_beams ? c->dbf_beams : c->cfg.n_lanes);
    if (!c->last_pc) return fail(c, RB200_ERR_ARG, "debug_fetch_pc: the last chain call kept its intermediate on chip (call rb200_set_debug_keep_pc(ctx, 1) first)");
    CK(c, cudaDeviceSynchronize());
    CK(c, cudaMemcpy(out_ri, c->last_pc + (size_t)cpi_in_chunk * cpi_cells, cpi_cells * sizeof(float2), cudaMemcpyDeviceToHost));
    return RB200_OK;
}

extern "C" int rb200_last_device_ms(const rb200_ctx* c, float* ms) {
    if (!c || !ms || !c->have_timing) return RB200_ERR_ARG;
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return RB200_ERR_CUDA;
    return cudaEventElapsedTime(ms, c->ev0, c->ev1) == cudaSuccess ? RB200_OK : RB200_ERR_CUDA;
}

extern "C" int rb200_set_stage_timing(rb200_ctx* c, int on) {
    if (!c) return RB200_ERR_ARG;
    c->stage_timing = on != 0;
    c->stage_used = 0;
    c->stage_cpis.clear();
    return RB200_OK;
}

extern "C" int rb200_get_stage_ms(rb200_ctx* c, float ms[3], int* n_chunks, int* n_cpis) {
    if (!c || !ms) return RB200_ERR_ARG;
    cudaSetDevice(c->device);
    ms[0] = ms[1] = ms[2] = 0.f;
    int chunks = 0, cpis = 0;
    for (size_t i = 0; i + 3 < c->stage_used; i += 4) {
        if (cudaEventSynchronize(c->stage_events[i + 3]) != cudaSuccess) return RB200_ERR_CUDA;
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, c->stage_events[i + k], c->stage_events[i + k + 1]) != cudaSuccess) return RB200_ERR_CUDA;
            ms[k] += t;
        }
        cpis += c->stage_cpis[chunks];
        ++chunks;
    }
    if (n_chunks) *n_chunks = chunks;
    if (n_cpis) *n_cpis = cpis;
    c->stage_used = 0;
    c->stage_cpis.clear();
    return RB200_OK;
}

extern "C" int rb200_last_launch_count(const rb200_ctx* c, int* n) {
    if (!c || !n) return RB200_ERR_ARG;
    *n = c->launches;
    return RB200_OK;
}
