// radix.cuh -- register-resident small DFT butterflies (radix 2/4/8/16), forward and inverse.
//
// Building blocks of the fast-time (pulse compression) and slow-time (MTD) FFT kernels.  The
// reference delegates these transforms to MATLAB's fft/ifft (MP/fun_pulse_compression.m:19-22,
// MP/fun_Process_MTD.m:24); here they are explicit fp32 butterflies.  Compiles for host as well
// (tests/host/test_radix.cpp checks every butterfly against a naive DFT).
#pragma once

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define RB_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#endif

namespace rb {

#if defined(__CUDA_ARCH__) && !defined(RB_NO_PACKED_F32)
// sm_100a packed fp32: one FADD2 / FMUL2 / FFMA2 works on a 64-bit register pair, i.e. on one complex number.  The
// SASS operands take a half swap (.LO_HI), a per-half negation and a scalar broadcast for free, so a complex
// add/subtract, a rotation by +-i folded into the following add and a scale are ONE instruction and a complex multiply
// is TWO.  The chain is issue-bound on fp32, so this halves its dominant instruction class.
// Results are identical to the scalar forms (same IEEE operations, same fused multiply-adds).
#define RB_PACKED_F32 1
__device__ __forceinline__ unsigned long long rb_pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long rb_pk(float2 a) { return rb_pk(a.x, a.y); }
__device__ __forceinline__ float2 rb_up(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long rb_add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long rb_sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long rb_mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long rb_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
RB_HD float2 cadd(float2 a, float2 b) { return rb_up(rb_add2(rb_pk(a), rb_pk(b))); }
RB_HD float2 csub(float2 a, float2 b) { return rb_up(rb_sub2(rb_pk(a), rb_pk(b))); }
// (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
RB_HD float2 cmul(float2 a, float2 b) {
    return rb_up(rb_fma2(rb_pk(a.x, a.x), rb_pk(b.x, b.y), rb_mul2(rb_pk(a.y, a.y), rb_pk(-b.y, b.x))));
}
// a * conj(b) = a.x * (b.x, -b.y) + a.y * (b.y, b.x)
RB_HD float2 cmulc(float2 a, float2 b) {
    return rb_up(rb_fma2(rb_pk(a.x, a.x), rb_pk(b.x, -b.y), rb_mul2(rb_pk(a.y, a.y), rb_pk(b.y, b.x))));
}
RB_HD float2 cscale(float2 a, float s) { return rb_up(rb_mul2(rb_pk(a), rb_pk(s, s))); }
template <int SIGN> RB_HD float2 mul_i(float2 a) {
    return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// exp(SIGN*i*pi/4) * a = h * (a + SIGN*i*a)
template <int SIGN> RB_HD float2 mul_w8_1(float2 a) {
    const float h = 0.70710678118654752440f;
    return rb_up(rb_mul2(rb_add2(rb_pk(a), rb_pk(mul_i<SIGN>(a))), rb_pk(h, h)));
}
// exp(SIGN*i*3pi/4) * a = h * (SIGN*i*a - a)
template <int SIGN> RB_HD float2 mul_w8_3(float2 a) {
    const float h = 0.70710678118654752440f;
    return rb_up(rb_mul2(rb_sub2(rb_pk(mul_i<SIGN>(a)), rb_pk(a)), rb_pk(h, h)));
}
#else
RB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
RB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
RB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
RB_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
RB_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// multiply by -i (forward, SIGN=-1) or +i (inverse, SIGN=+1)
template <int SIGN> RB_HD float2 mul_i(float2 a) {
    return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// multiply by exp(SIGN * i*pi/4) = (1 + SIGN*i)/sqrt2
template <int SIGN> RB_HD float2 mul_w8_1(float2 a) {
    const float h = 0.70710678118654752440f;
    return SIGN < 0 ? make_float2((a.x + a.y) * h, (a.y - a.x) * h) : make_float2((a.x - a.y) * h, (a.y + a.x) * h);
}
// multiply by exp(SIGN * i*3pi/4) = (-1 + SIGN*i)/sqrt2
template <int SIGN> RB_HD float2 mul_w8_3(float2 a) {
    const float h = 0.70710678118654752440f;
    return SIGN < 0 ? make_float2((a.y - a.x) * h, -(a.x + a.y) * h) : make_float2(-(a.x + a.y) * h, (a.x - a.y) * h);
}
#endif

// X[k] = sum_j x[j] exp(SIGN*2*pi*i*j*k/R), natural order in, natural order out, in place.
template <int SIGN> RB_HD void dft2(float2& a, float2& b) {
    float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <int SIGN> RB_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i<SIGN>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

template <int SIGN> RB_HD void dft8(float2 (&v)[8]) {
    // even / odd radix-4 sub-transforms, then combine with w8^k
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<SIGN>(e0, e1, e2, e3);
    dft4<SIGN>(o0, o1, o2, o3);
    o1 = mul_w8_1<SIGN>(o1);
    o2 = mul_i<SIGN>(o2);
    o3 = mul_w8_3<SIGN>(o3);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <int SIGN> RB_HD void dft16(float2 (&v)[16]) {
    // 4 x radix-4 over j = j0 + 4*j1 (sum over j1), twiddle w16^(j0*k0), 4 x radix-4 over j0.
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos/sin(pi/8)
    const float h = 0.70710678118654752440f;
#pragma unroll
    for (int j0 = 0; j0 < 4; ++j0) dft4<SIGN>(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);
    // after: v[j0 + 4*k0] = sum_j1 x[j0+4 j1] w4^(j1 k0).  Twiddle by w16^(j0*k0), SIGN-dependent.
    const float sg = SIGN < 0 ? -1.f : 1.f;
    // j0=1: k0=1,2,3 -> angles 1,2,3 (x pi/8)
    v[1 + 4]  = cmul(v[1 + 4],  make_float2(c1, sg * s1));
    v[1 + 8]  = cmul(v[1 + 8],  make_float2(h, sg * h));
    v[1 + 12] = cmul(v[1 + 12], make_float2(s1, sg * c1));
    // j0=2: angles 2,4,6
    v[2 + 4]  = cmul(v[2 + 4],  make_float2(h, sg * h));
    v[2 + 8]  = mul_i<SIGN>(v[2 + 8]);
    v[2 + 12] = cmul(v[2 + 12], make_float2(-h, sg * h));
    // j0=3: angles 3,6,9
    v[3 + 4]  = cmul(v[3 + 4],  make_float2(s1, sg * c1));
    v[3 + 8]  = cmul(v[3 + 8],  make_float2(-h, sg * h));
    v[3 + 12] = cmul(v[3 + 12], make_float2(-c1, -sg * s1));
    // second pass: for each k0, radix-4 over j0 -> k1; X[k0 + 4*k1]
#pragma unroll
    for (int k0 = 0; k0 < 4; ++k0) dft4<SIGN>(v[4 * k0 + 0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3]);
    // now v[4*k0 + k1] = X[k0 + 4*k1]; transpose to natural order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            float2 t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
}

template <int R, int SIGN> struct Dft;
template <int SIGN> struct Dft<2, SIGN> { RB_HD static void run(float2 (&v)[2]) { dft2<SIGN>(v[0], v[1]); } };
template <int SIGN> struct Dft<4, SIGN> { RB_HD static void run(float2 (&v)[4]) { dft4<SIGN>(v[0], v[1], v[2], v[3]); } };
template <int SIGN> struct Dft<8, SIGN> { RB_HD static void run(float2 (&v)[8]) { dft8<SIGN>(v); } };
template <int SIGN> struct Dft<16, SIGN> { RB_HD static void run(float2 (&v)[16]) { dft16<SIGN>(v); } };

}  // namespace rb
