// fp32_rate.cu -- measures the issue rate of scalar and packed fp32 instructions on one GPU (warp-instructions per clock
// per SM).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o fp32_rate fp32_rate.cu ; run on the GPU box.
// Used to decide whether the FFT butterflies should use FADD2/FFMA2.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, float a, float b) {
    float x[NACC], y[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { x[i] = a + i + threadIdx.x; y[i] = b - i; }
    unsigned long long pa, pb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(a));
    unsigned long long p[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x[i]), "f"(y[i]));
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) { x[i] = x[i] + y[i]; y[i] = y[i] + a; }                                 // 2 FADD (reg+reg)
            if (MODE == 1) { x[i] = fmaf(x[i], a, y[i]); y[i] = fmaf(y[i], b, x[i]); }              // 2 FFMA 3-reg
            if (MODE == 2) { x[i] = fmaf(x[i], 1.0001f, 0.5f); y[i] = fmaf(y[i], 0.9999f, 0.25f); } // 2 FFMA imm
            if (MODE == 3) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb)); }
            if (MODE == 4) { asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb)); asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pa)); }
            if (MODE == 5) { x[i] = x[i] * a; y[i] = y[i] * b; }                                     // 2 FMUL
            if (MODE == 6) { x[i] = x[i] + y[i]; asm("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa)); }   // 1 FADD + 1 FADD2
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
        s += x[i] + y[i] + lo + hi;
    }
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
static void run(const char* name, int sms, int clock_khz) {
    float* d;
    cudaMalloc(&d, 4);
    const int ctas = sms * 8;        // 2048 threads per SM
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<MODE><<<ctas, 256>>>(d, 1.5f, 0.25f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_kernel<MODE><<<ctas, 256>>>(d, 1.5f, 0.25f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)ctas * 8 * ITERS * NACC * 2;      // warp-instructions of the measured kind
    const double cycles = ms * 1e-3 * clock_khz * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, winst / cycles / sms, clock_khz / 1000);
    cudaFree(d);
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, clock attr %d kHz\n", pr.name, pr.multiProcessorCount, khz);
    run<0>("FADD (2 per step)", pr.multiProcessorCount, khz);
    run<1>("FFMA 3-reg", pr.multiProcessorCount, khz);
    run<2>("FFMA imm", pr.multiProcessorCount, khz);
    run<5>("FMUL", pr.multiProcessorCount, khz);
    run<3>("FADD2 (packed)", pr.multiProcessorCount, khz);
    run<4>("FFMA2 (packed)", pr.multiProcessorCount, khz);
    run<6>("FADD + FADD2 mixed", pr.multiProcessorCount, khz);
    return 0;
}
