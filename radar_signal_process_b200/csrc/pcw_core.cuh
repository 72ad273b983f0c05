// pcw_core.cuh -- the warp-private pulse-compression transform shared by pcw_kernel.cu (K1, wire format) and
// onepass_kernel.cu (single-pass chain): 256-point overlap-save fast correlation of TWO lines per thread
// (MP/fun_pulse_compression.m:16-22: fft -> conj reference spectrum -> ifft), 16 threads per line pair.
#pragma once
#include "common.cuh"
#include "radix.cuh"

namespace rb {

// offset-binary int16 pair -> (float I, float Q), exact, without I2F: 0x4B400000 | u is the float 12582912 + u, u = s + 32768
__device__ __forceinline__ float2 unpack_ob(unsigned w) {
    const unsigned fi = __byte_perm(w, 0x4B400000u, 0x7610);
    const unsigned fq = __byte_perm(w, 0x4B400000u, 0x7632);
    return csub(make_float2(__uint_as_float(fi), __uint_as_float(fq)), make_float2(12615680.f, 12615680.f));
}

// complex products with a factor held as two scalars: a * (bx + i by) = a * bx + (-a.y, a.x) * by -- on sm_100a one FMUL2
// and one FFMA2 (the scalar is a broadcast operand, the swap / half negation are operand modifiers)
#if defined(RB_PACKED_F32)
__device__ __forceinline__ float2 cmul_s(float2 a, float bx, float by) {
    return rb_up(rb_fma2(rb_pk(-a.y, a.x), rb_pk(by, by), rb_mul2(rb_pk(a.x, a.y), rb_pk(bx, bx))));
}
__device__ __forceinline__ float2 cmulc_s(float2 a, float bx, float by) {            // a * conj(b) = a * bx + (a.y, -a.x) * by
    return rb_up(rb_fma2(rb_pk(a.y, -a.x), rb_pk(by, by), rb_mul2(rb_pk(a.x, a.y), rb_pk(bx, bx))));
}
#else       // host compilation pass: the same products in scalar form
__device__ __forceinline__ float2 cmul_s(float2 a, float bx, float by) { return cmul(a, make_float2(bx, by)); }
__device__ __forceinline__ float2 cmulc_s(float2 a, float bx, float by) { return cmulc(a, make_float2(bx, by)); }
#endif


// two's-complement int16 pair (wire format, FrameDataRead_xzr.m:154-156) -> (float I, float Q), exact, without I2F
__device__ __forceinline__ float2 unpack_tc(unsigned w) { return unpack_ob(w ^ 0x80008000u); }

constexpr int kPcwRowC = 272;      // complex slots per exchange row: the 16 x 16 matrix at a pitch of 17 slots

// In: a[j], b[j] = x[n1 + 16 j] of the thread's two lines.  Out: a[j], b[j] = y[n1 + 16 j], y = ifft(fft(x) .* H) with
// H = conj(FFT(taps)) * scale / 256 given as h_blk[16 j + n1] = bin n1 + 16 j, tw = w256^(n1 q) (PcTwiddles).
// Radix-16 x 16: forward DIF, spectrum product in digit-reversed order, inverse DIT; the two exchanges go through the
// line's own row (wA/wB = row + n1, rA/rB = row + 17 n1): element (r, c) of the 16 x 16 matrix sits at slot 17 r + c,
// writers fill 16 consecutive slots per instruction, readers hit 16 different bank pairs, every address is base +
// immediate.  Only __syncwarp between the phases: a row is private to the 16 threads of its line pair.
// the thread's twiddles w256^(n1 q), q = 1..15, as scalar pairs
struct PcTwiddles {
    float x[15], y[15];
    __device__ __forceinline__ void load(const float2* tw_sm, int n1) {
#pragma unroll
        for (int q = 1; q < 16; ++q) {
            const float2 w = tw_sm[q * 16 + n1];
            x[q - 1] = w.x;
            y[q - 1] = w.y;
        }
    }
};

__device__ __forceinline__ void pc_pair_transform(float2 (&a)[16], float2 (&b)[16], const PcTwiddles& tw, const float2* h_blk, int n1,
                                                  float2* wA, float2* wB, const float2* rA, const float2* rB) {
    const float (&twx)[15] = tw.x;
    const float (&twy)[15] = tw.y;
    Dft<16, -1>::run(a);
    Dft<16, -1>::run(b);
#pragma unroll
    for (int q = 1; q < 16; ++q) {
        a[q] = cmul_s(a[q], twx[q - 1], twy[q - 1]);
        b[q] = cmul_s(b[q], twx[q - 1], twy[q - 1]);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        wA[17 * q] = a[q];
        wB[17 * q] = b[q];
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        a[m] = rA[m];
        b[m] = rB[m];
    }
    Dft<16, -1>::run(a);                                            // a[j] = X[n1 + 16 j]
    float2 hv[16];                                                  // spectrum values fetched under the second line's butterflies
#pragma unroll
    for (int j = 0; j < 16; ++j) hv[j] = h_blk[16 * j + n1];
    Dft<16, -1>::run(b);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        a[j] = cmul_s(a[j], hv[j].x, hv[j].y);
        b[j] = cmul_s(b[j], hv[j].x, hv[j].y);
    }
    Dft<16, +1>::run(a);
    Dft<16, +1>::run(b);
#pragma unroll
    for (int m = 1; m < 16; ++m) {
        a[m] = cmulc_s(a[m], twx[m - 1], twy[m - 1]);
        b[m] = cmulc_s(b[m], twx[m - 1], twy[m - 1]);
    }
    __syncwarp();                                                   // the first exchange has been read by everyone
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        wA[17 * m] = a[m];
        wB[17 * m] = b[m];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        a[q] = rA[q];
        b[q] = rB[q];
    }
    Dft<16, +1>::run(a);                                            // a[j] = y[n1 + 16 j]
    Dft<16, +1>::run(b);
}

}  // namespace rb
