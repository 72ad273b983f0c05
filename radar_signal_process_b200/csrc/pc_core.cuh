// pc_core.cuh -- device building blocks of the pulse-compression kernels (shared by pc_kernels.cu and
// chain64_kernel.cu): stage-major twiddle addressing, the forward-DIF / spectrum / inverse-DIT core on the R
// operands of one thread, and the mbarrier / TMA bulk-copy PTX wrappers.
#pragma once
#include "common.cuh"
#include "radix.cuh"

namespace rb {

// offset (in float2) of the stage-s twiddle block inside PcParams::tw:  T_s[k*st_s + q] = w_NT^(q*k*R^s),
// st_s = NT / R^(s+1).  Blocks exist for s = 0 .. S-2 (the last stage has no twiddles).
template <int R, int S> __host__ __device__ constexpr int tw_block_off(int s) {
    int off = 0;
    for (int i = 0; i < s; ++i) off += R * (ipow(R, S) / ipow(R, i + 1));
    return off;
}

template <int R, int S, int LT> struct PcOcc { static constexpr int min_blocks = (LT * (ipow(R, S) / R) <= 256) ? 3 : 1; };

// One forward DIF stage s >= 1: exchange through shared memory (write at the positions of stage s-1, read at
// the positions of stage s), butterfly, twiddle.  All strides are template constants.
// Table loads: the persistent kernels keep the twiddle and spectrum tables in shared memory (TAB_SMEM), the
// one-tile-per-CTA kernels read them through the read-only global path.
template <bool TAB_SMEM> __device__ __forceinline__ float2 tab_ld(const float2* ptr) { return TAB_SMEM ? *ptr : __ldg(ptr); }
template <bool TAB_SMEM> __device__ __forceinline__ float4 tab_ld4(const float4* ptr) { return TAB_SMEM ? *ptr : __ldg(ptr); }

template <int R, int S, int s, bool TAB_SMEM>
__device__ __forceinline__ void pc_fwd_stage(float2 (&v)[R], float2* line_sm, const float2* __restrict__ twtab, int u) {
    constexpr int NT = ipow(R, S);
    constexpr int sp = NT / ipow(R, s);            // stride of stage s-1
    constexpr int st = NT / ipow(R, s + 1);        // stride of stage s
    constexpr int twoff = tw_block_off<R, S>(s);
    const int basep = (u / sp) * sp * R + (u % sp);
#pragma unroll
    for (int k = 0; k < R; ++k) line_sm[basep + k * sp] = v[k];
    __syncthreads();
    const int q = u % st;
    const int base = (u / st) * st * R + q;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = line_sm[base + j * st];
    Dft<R, -1>::run(v);
    if (st > 1) {
        const float2* tw = twtab + twoff + q;
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmul(v[k], tab_ld<TAB_SMEM>(tw + k * st));
    }
}

// One inverse DIT stage s >= 1: conjugate twiddle, butterfly, exchange towards stage s-1.  When s-1 == 0 the
// read side is re-mapped butterfly-fastest (out_lane, out_u) so the final stores coalesce along range.
template <int R, int S, int s, int LT, bool TAB_SMEM>
__device__ __forceinline__ void pc_inv_stage(float2 (&v)[R], float2* sm, const float2* __restrict__ twtab, int t, int lane, int u,
                                             int& out_lane, int& out_u) {
    constexpr int NT = ipow(R, S);
    constexpr int NB = NT / R;
    constexpr int LS = NT + 1;
    constexpr int st = NT / ipow(R, s + 1);
    constexpr int sn = NT / ipow(R, s);            // stride of stage s-1
    constexpr int twoff = tw_block_off<R, S>(s);
    float2* line_sm = sm + lane * LS;
    const int q = u % st;
    if (st > 1) {
        const float2* tw = twtab + twoff + q;
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], tab_ld<TAB_SMEM>(tw + k * st));
    }
    Dft<R, +1>::run(v);
    const int base = (u / st) * st * R + q;
#pragma unroll
    for (int j = 0; j < R; ++j) line_sm[base + j * st] = v[j];
    __syncthreads();
    if (s - 1 == 0) {
        out_lane = t / NB;
        out_u = t % NB;
        const float2* src = sm + out_lane * LS + out_u;
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = src[k * sn];
    } else {
        const int basen = (u / sn) * sn * R + (u % sn);
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = line_sm[basen + k * sn];
    }
}

// Forward DIF -> spectrum multiply -> inverse DIT on the R operands of one thread (see file header).
// On return v[j] holds lag out_u + j*NB of line out_lane (butterfly-fastest mapping for coalesced stores).
// twtab / hseg: twiddle table and the segment's digit-reversed spectrum (shared memory if TAB_SMEM, else global).
template <int R, int S, int LT, bool TAB_SMEM = false>
__device__ __forceinline__ void pc_fft_core(float2 (&v)[R], float2* sm, const float2* __restrict__ twtab, const float2* __restrict__ hseg,
                                            int t, int lane, int u, int& out_lane_r, int& out_u_r) {
    static_assert(S >= 1 && S <= 3, "1..3 stages");
    constexpr int NT = ipow(R, S);
    constexpr int NB = NT / R;
    constexpr int LS = NT + 1;
    float2* line_sm = sm + lane * LS;

    // ---- forward DIF ----
    Dft<R, -1>::run(v);
    if (S > 1) {
        const float2* tw = twtab + u;      // stage-0 block: T[k*NB + u]
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmul(v[k], tab_ld<TAB_SMEM>(tw + k * NB));
    }
    if (S >= 2) pc_fwd_stage<R, S, 1, TAB_SMEM>(v, line_sm, twtab, u);
    if (S >= 3) pc_fwd_stage<R, S, (S >= 3 ? 2 : 1), TAB_SMEM>(v, line_sm, twtab, u);

    // ---- reference spectrum (digit-reversed order == this thread's positions u*R .. u*R+R-1) ----
    {
        const float4* hp = reinterpret_cast<const float4*>(hseg + u * R);
#pragma unroll
        for (int k = 0; k < R; k += 2) {
            const float4 h = tab_ld4<TAB_SMEM>(hp + k / 2);
            v[k] = cmul(v[k], make_float2(h.x, h.y));
            v[k + 1] = cmul(v[k + 1], make_float2(h.z, h.w));
        }
    }

    // ---- inverse DIT ----
    int out_lane = lane, out_u = u;
    if (S >= 3) pc_inv_stage<R, S, (S >= 3 ? 2 : 1), LT, TAB_SMEM>(v, sm, twtab, t, lane, u, out_lane, out_u);
    if (S >= 2) pc_inv_stage<R, S, 1, LT, TAB_SMEM>(v, sm, twtab, t, lane, u, out_lane, out_u);
    if (S > 1) {
        // stage 0 (operands fetched with the store mapping)
        const float2* tw = twtab + out_u;
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], tab_ld<TAB_SMEM>(tw + k * NB));
    }
    Dft<R, +1>::run(v);
    out_lane_r = out_lane;
    out_u_r = out_u;
}

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

}  // namespace rb
