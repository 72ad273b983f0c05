// fp32x2_forms.cu -- issue rate of the packed fp32 operand forms used by pcw_core.cuh (warp-instructions per clock per SM):
// plain FADD2, FADD2 with a swapped/negated operand, FMUL2 with a scalar-broadcast operand, FFMA2 with both, and the two-
// instruction complex product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o fp32x2_forms fp32x2_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 16384
#define NACC 8

__device__ __forceinline__ unsigned long long pk(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float sa, float sb, int nwarps_per_cta) {
    if ((int)(threadIdx.x >> 5) >= nwarps_per_cta) return;
    float x[NACC], y[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { x[i] = sa + i + threadIdx.x; y[i] = sb - i; }
    float bx = sa * 0.999f, by = sb * 1.001f;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            unsigned long long a = pk(x[i], y[i]);
            if (MODE == 0) { a = add2(a, pk(bx, by)); a = add2(a, pk(by, bx)); }                         // 2 plain FADD2
            if (MODE == 1) { a = add2(a, pk(-y[i], x[i])); a = add2(a, pk(y[i], -x[i])); }               // 2 FADD2, swapped + half-negated operand
            if (MODE == 2) { a = mul2(a, pk(bx, bx)); a = mul2(a, pk(by, by)); }                         // 2 FMUL2, scalar broadcast
            if (MODE == 3) { a = fma2(pk(-y[i], x[i]), pk(by, by), a); a = fma2(pk(y[i], -x[i]), pk(bx, bx), a); }   // 2 FFMA2 swap/neg + broadcast
            if (MODE == 5) { a = fma2(a, pk(bx, by), pk(by, bx)); a = fma2(a, pk(by, bx), pk(bx, by)); }             // 2 FFMA2, three plain register pairs
            if (MODE == 6) { a = fma2(a, pk(bx, bx), pk(by, bx)); a = fma2(a, pk(by, by), pk(bx, by)); }             // 2 FFMA2, broadcast operand only
            if (MODE == 7) { a = fma2(pk(-y[i], x[i]), pk(bx, by), a); a = fma2(pk(y[i], -x[i]), pk(by, bx), a); }   // 2 FFMA2, swap/neg operand only
            if (MODE == 8) { a = mul2(pk(-y[i], x[i]), pk(by, by)); a = mul2(pk(y[i], -x[i]), pk(bx, bx)); }          // 2 FMUL2, swap/neg + broadcast
            if (MODE == 9) { unsigned long long t = mul2(pk(-y[i], x[i]), pk(by, by)); a = fma2(a, pk(bx, bx), t); }  // complex product, swap on the FMUL2
            if (MODE == 4) { unsigned long long r = mul2(a, pk(bx, bx)); a = fma2(pk(-y[i], x[i]), pk(by, by), r); } // complex product (FMUL2 + FFMA2)
            up(a, x[i], y[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += x[i] + y[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
static void run(const char* name, int sms, int khz, int warps_per_sm) {
    float* d;
    cudaMalloc(&d, 4);
    const int ctas = sms * 2;               // 2 CTAs per SM, warps_per_sm / 2 active warps each
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(d, 1.5f, 0.25f, warps_per_sm / 2);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(d, 1.5f, 0.25f, warps_per_sm / 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)sms * warps_per_sm * ITERS * NACC * 2;
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-44s %2d warps/SM %8.3f ms  %5.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, warps_per_sm, ms, winst / cycles / sms, khz / 1000);
    cudaFree(d);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, clock attr %d kHz\n", pr.name, pr.multiProcessorCount, khz);
    for (int w : {16}) {
        run<0>("FADD2 plain", pr.multiProcessorCount, khz, w);
        run<1>("FADD2 swapped/half-negated operand", pr.multiProcessorCount, khz, w);
        run<2>("FMUL2 scalar-broadcast operand", pr.multiProcessorCount, khz, w);
        run<3>("FFMA2 swap/neg + broadcast", pr.multiProcessorCount, khz, w);
        run<4>("complex product FMUL2 + FFMA2(swap)", pr.multiProcessorCount, khz, w);
        run<5>("FFMA2 three plain pairs", pr.multiProcessorCount, khz, w);
        run<6>("FFMA2 broadcast only", pr.multiProcessorCount, khz, w);
        run<7>("FFMA2 swap/neg only", pr.multiProcessorCount, khz, w);
        run<8>("FMUL2 swap/neg + broadcast", pr.multiProcessorCount, khz, w);
        run<9>("complex product FMUL2(swap) + FFMA2", pr.multiProcessorCount, khz, w);
    }
    return 0;
}
