// dft16_rate.cu -- how fast do the register-resident radix-16 butterflies of radix.cuh run when nothing else is in the
// way (no shared memory, no barriers, no global traffic)?  Reports the time per butterfly per SM and the implied fp32
// lane rate, for 8 / 16 / 24 warps per SM, so that the pulse-compression kernel's own efficiency can be judged.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a [-DRB_NO_PACKED_F32] -I../../radar_signal_process_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "radix.cuh"

using namespace rb;

#define ITERS 2000

__global__ void __launch_bounds__(256) dft16_loop(float2* out, float seed) {
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(seed + i + threadIdx.x, seed - i);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        Dft<16, -1>::run(v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = cscale(v[i], 0.25f);        // keeps magnitudes bounded (16 packed / 32 scalar multiplies)
    }
    float2 s = v[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) s = cadd(s, v[i]);
    if (s.x == 1234.5f) out[0] = s;
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float2* d;
    cudaMalloc(&d, 8);
    const int sms = pr.multiProcessorCount;
#ifdef RB_NO_PACKED_F32
    printf("scalar fp32 butterflies\n");
#else
    printf("packed fp32x2 butterflies\n");
#endif
    for (int ctas = 1; ctas <= 4; ++ctas) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        dft16_loop<<<sms * ctas, 256>>>(d, 1.f);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        dft16_loop<<<sms * ctas, 256>>>(d, 1.f);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double cycles = ms * 1e-3 * khz * 1e3;
        const double bf_per_sm = (double)ctas * 256 * ITERS;                 // butterflies (+16 scalings) per SM
        // 144 real adds + 24 real multiplies-ish per dft16 (168 flops) + 32 for the scaling = 200 lane-ops per butterfly
        printf("%d CTAs/SM (%2d warps): %7.3f ms, %6.1f cycles per warp-butterfly per SM, %.1f fp32 lane-ops/clk/SM (peak 128)\n", ctas,
               ctas * 8, ms, cycles / (bf_per_sm / 32), bf_per_sm * 200.0 / cycles);
    }
    return 0;
}
