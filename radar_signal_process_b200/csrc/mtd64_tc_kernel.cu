// mtd64_tc_kernel.cu -- EXPERIMENT (RB200_MTD_TC=1): the 64-point slow-time transform of MP/fun_Process_MTD.m:20-26 (Kaiser
// window, fft along the PRT axis, fftshift, abs) and the row zeroing of MP/fun_0v_pressing.m:4-6 as a DFT-by-GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), to be measured against the butterfly kernels
// (mtd64_kernel.cu / mtd_kernels.cu).  BASELINE.json's north star allows tensor cores only if such a variant beats the
// butterflies in ncu; profiles/README.md records the outcome.
//
// Formulation.  For one (CPI, lane) slab and a tile of 32 range cells n, with E[k][p] = keep[k] w[p] exp(-2 pi i f p / 64),
// f = (k + 32) mod 64 (fftshift folded into the row order, window and zero-velocity rows folded into the matrix):
//     D[k][2n]   = sum_p  Er[k][p] Xr[p][n] + Ei[k][p] (-Xi[p][n])      = Re Y[k][n]
//     D[k][2n+1] = sum_p  Er[k][p] Xi[p][n] + Ei[k][p]   Xr[p][n]       = Im Y[k][n]
// i.e. one real GEMM D(64 x 64) = A(64 x 128) B(128 x 64) per tile with both parts of an output in the SAME TMEM lane, so
// the epilogue thread forms |Y| without a cross-lane exchange.  fp32 accuracy is approximated with a two-term bf16 split
// of both operands, x = hi + lo, and three products hi*hi + lo*hi + hi*lo (relative error ~2^-16 per element, inside the
// 1e-4 RDM gate): A' = [Ahi | Ahi | Alo] (64 x 384, built once on the host), B' = [Bhi ; Blo ; Bhi] (384 x 64, converted
// from the fp32 pulse-compressed samples by the CTA), 24 K-steps of tcgen05.mma.kind::f16 M64 N64 K16 per tile.
// Operands sit in shared memory in the canonical K-major no-swizzle layout: 8-element (16-byte) K chunks, chunk c of row r
// at c * 1024 + r * 16 (core matrices of 8 rows x 16 bytes, SBO = 128 bytes, LBO = 1024 bytes).
#include "common.cuh"
#include "kernels.h"
#include <cuda_bf16.h>
#include <vector>
#include <cmath>

namespace rb {

namespace tc {
constexpr int kP = 64;                 // PRTs = DFT length = GEMM M
constexpr int kNr = 32;                // range cells per tile
constexpr int kN = 2 * kNr;            // GEMM N: (re, im) columns
constexpr int kKterm = 2 * kP;         // (p, component)
constexpr int kK = 3 * kKterm;         // three split products
constexpr int kChunks = kK / 8;        // 16-byte K chunks
constexpr int kChunkBytesA = kP * 16;  // 1024
constexpr int kChunkBytesB = kN * 16;  // 1024
constexpr int kABytes = kChunks * kChunkBytesA;   // 49 152
constexpr int kBBytes = kChunks * kChunkBytesB;   // 49 152
constexpr int kXBytes = kP * kNr * 8;             // 16 384: fp32 complex staging [p][n]
constexpr int kSmem = kABytes + kBBytes + kXBytes;
constexpr int kThreads = 128;
constexpr int kTmemCols = 64;
}  // namespace tc

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: start address, leading byte offset (between the two 16-byte K chunks of one MMA), stride byte offset
// (between 8-row core matrices), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

__global__ void __launch_bounds__(tc::kThreads, 2)
mtd64_tc_kernel(const float2* __restrict__ in, float* __restrict__ out, const uint4* __restrict__ a_mat, int in_ld, int out_ld, int cols,
                int tiles_per_slab, int n_items) {
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + kABytes;
    float2* sX = reinterpret_cast<float2*>(smem + kABytes + kBBytes);
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_sm;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_sm)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = t; i < kABytes / 16; i += kThreads) reinterpret_cast<uint4*>(sA)[i] = __ldg(a_mat + i);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sm;

    // instruction descriptor: D fp32, A/B bf16, both K-major, N = 64, M = 64
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kP >> 4) << 24);
    const uint32_t sA_u = tc_smem_u32(sA), sB_u = tc_smem_u32(sB);

    // the fp32 tile [p][n] of an item travels through 16 registers per thread (a warp reads 256 contiguous bytes of one PRT
    // line per load); the loads of the NEXT item are issued before the current one is converted, multiplied and stored
    float2 xr[kP * kNr / kThreads];
    auto fetch = [&](int item) {
        const int slab = item / tiles_per_slab;
        const int r0 = (item - slab * tiles_per_slab) * kNr;
        const float2* src = in + (size_t)slab * kP * in_ld + r0;
#pragma unroll
        for (int j = 0; j < kP * kNr / kThreads; ++j) {
            const int i = t + j * kThreads;
            const int p = i >> 5, n = i & 31;
            xr[j] = (r0 + n < cols) ? __ldg(src + (size_t)p * in_ld + n) : make_float2(0.f, 0.f);
        }
    };
    if ((int)blockIdx.x < n_items) fetch(blockIdx.x);
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int slab = item / tiles_per_slab;
        const int r0 = (item - slab * tiles_per_slab) * kNr;
#pragma unroll
        for (int j = 0; j < kP * kNr / kThreads; ++j) sX[t + j * kThreads] = xr[j];
        __syncthreads();
        if (item + (int)gridDim.x < n_items) fetch(item + gridDim.x);
        // ---- convert: thread = (range cell n, group g of 8 PRTs); writes whole 16-byte K chunks of columns 2n and 2n + 1
        for (int task = t; task < kNr * 8; task += kThreads) {
            const int n = task & 31, g = task >> 5;
            uint32_t rh[4], rl[4], ih[4], il[4], nih[4], nil[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 x0 = sX[(8 * g + 2 * q) * kNr + n];
                const float2 x1 = sX[(8 * g + 2 * q + 1) * kNr + n];
                const __nv_bfloat16 r0h = __float2bfloat16_rn(x0.x), r1h = __float2bfloat16_rn(x1.x);
                const __nv_bfloat16 i0h = __float2bfloat16_rn(x0.y), i1h = __float2bfloat16_rn(x1.y);
                const __nv_bfloat16 r0l = __float2bfloat16_rn(x0.x - __bfloat162float(r0h)), r1l = __float2bfloat16_rn(x1.x - __bfloat162float(r1h));
                const __nv_bfloat16 i0l = __float2bfloat16_rn(x0.y - __bfloat162float(i0h)), i1l = __float2bfloat16_rn(x1.y - __bfloat162float(i1h));
                rh[q] = pack_bf16(r0h, r1h);
                rl[q] = pack_bf16(r0l, r1l);
                ih[q] = pack_bf16(i0h, i1h);
                il[q] = pack_bf16(i0l, i1l);
                nih[q] = ih[q] ^ 0x80008000u;
                nil[q] = il[q] ^ 0x80008000u;
            }
            // column 2n   (Re): component 0 = Xr, component 1 = -Xi ; column 2n+1 (Im): component 0 = Xi, component 1 = Xr
            // K index = term * 128 + component * 64 + p -> chunk = term * 16 + component * 8 + g ; terms: hi, lo, hi
            unsigned char* colRe = sB + (2 * n) * 16;
            unsigned char* colIm = sB + (2 * n + 1) * 16;
            const uint4 Rh = make_uint4(rh[0], rh[1], rh[2], rh[3]), Rl = make_uint4(rl[0], rl[1], rl[2], rl[3]);
            const uint4 Ih = make_uint4(ih[0], ih[1], ih[2], ih[3]), Il = make_uint4(il[0], il[1], il[2], il[3]);
            const uint4 NIh = make_uint4(nih[0], nih[1], nih[2], nih[3]), NIl = make_uint4(nil[0], nil[1], nil[2], nil[3]);
            *reinterpret_cast<uint4*>(colRe + (0 * 16 + g) * kChunkBytesB) = Rh;
            *reinterpret_cast<uint4*>(colRe + (0 * 16 + 8 + g) * kChunkBytesB) = NIh;
            *reinterpret_cast<uint4*>(colRe + (1 * 16 + g) * kChunkBytesB) = Rl;
            *reinterpret_cast<uint4*>(colRe + (1 * 16 + 8 + g) * kChunkBytesB) = NIl;
            *reinterpret_cast<uint4*>(colRe + (2 * 16 + g) * kChunkBytesB) = Rh;
            *reinterpret_cast<uint4*>(colRe + (2 * 16 + 8 + g) * kChunkBytesB) = NIh;
            *reinterpret_cast<uint4*>(colIm + (0 * 16 + g) * kChunkBytesB) = Ih;
            *reinterpret_cast<uint4*>(colIm + (0 * 16 + 8 + g) * kChunkBytesB) = Rh;
            *reinterpret_cast<uint4*>(colIm + (1 * 16 + g) * kChunkBytesB) = Il;
            *reinterpret_cast<uint4*>(colIm + (1 * 16 + 8 + g) * kChunkBytesB) = Rl;
            *reinterpret_cast<uint4*>(colIm + (2 * 16 + g) * kChunkBytesB) = Ih;
            *reinterpret_cast<uint4*>(colIm + (2 * 16 + 8 + g) * kChunkBytesB) = Rh;
        }
        // generic-proxy writes of the operand tile -> visible to the tensor core (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int ks = 0; ks < kK / 16; ++ks) {
                const uint64_t da = tc_desc(sA_u + ks * 2 * kChunkBytesA, kChunkBytesA, 128);
                const uint64_t db = tc_desc(sB_u + ks * 2 * kChunkBytesB, kChunkBytesB, 128);
                const uint32_t acc = ks > 0 ? 1u : 0u;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
                    "}\n" ::"r"(tmem),
                    "l"(da), "l"(db), "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&mma_bar)) : "memory");
        }
        // ---- wait for the accumulator, |Y| = sqrt(Re^2 + Im^2), rows = output rows (fftshift and 0-v rows are in the matrix)
        {
            const uint32_t parity = (uint32_t)(it & 1);
            asm volatile(
                "{\n"
                ".reg .pred P1;\n"
                "TC_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                "@P1 bra TC_DONE;\n"
                "bra TC_WAIT;\n"
                "TC_DONE:\n"
                "}" ::"r"(tc_smem_u32(&mma_bar)),
                "r"(parity)
                : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            // M = 64 accumulator: row 16 w + j sits in TMEM lane 32 w + j (j < 16); a warp reads its own 32 lanes
            uint32_t v[64];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
                  "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
                  "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
                  "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (lane < 16) {
                const int k = 16 * warp + lane;
                float* o = out + ((size_t)slab * kP + k) * out_ld + r0;
#pragma unroll
                for (int n = 0; n < kNr; n += 4) {
                    float m[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float re = __uint_as_float(v[2 * (n + j)]), im = __uint_as_float(v[2 * (n + j) + 1]);
                        m[j] = sqrtf(re * re + im * im);
                    }
                    if (r0 + n + 3 < cols && (out_ld & 3) == 0) {
                        *reinterpret_cast<float4*>(o + n) = make_float4(m[0], m[1], m[2], m[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (r0 + n + j < cols) o[n + j] = m[j];
                    }
                }
            }
        }
        // the accumulator and the operand tiles are reused by the next item
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

// ---------------------------------------------------------------------------------------------
static uint16_t host_bf16(float f) {       // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(r >> 16);
}
static float host_bf16_to_float(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// A' = [Ahi | Ahi | Alo], A = [Er | Ei] with E[k][p] = keep[k] w[p] exp(-2 pi i ((k + 32) mod 64) p / 64), in the shared-memory
// layout of the kernel (chunk c of row k at c * 1024 + k * 16 bytes)
void mtd64_tc_build_matrix(const float* window, int zv_lo, int zv_hi, std::vector<uint16_t>& a) {
    using namespace tc;
    a.assign(kABytes / 2, 0);
    for (int k = 0; k < kP; ++k) {
        const bool zero = k >= zv_lo && k <= zv_hi;
        const int f = (k + kP / 2) & (kP - 1);
        for (int p = 0; p < kP; ++p) {
            const double th = 2.0 * M_PI * (double)((f * p) & (kP - 1)) / kP;
            const float er = zero ? 0.f : (float)((double)window[p] * std::cos(th));
            const float ei = zero ? 0.f : (float)(-(double)window[p] * std::sin(th));
            const float comp[2] = {er, ei};
            for (int c = 0; c < 2; ++c) {
                const uint16_t hi = host_bf16(comp[c]);
                const uint16_t lo = host_bf16(comp[c] - host_bf16_to_float(hi));
                const uint16_t term[3] = {hi, hi, lo};
                for (int tm = 0; tm < 3; ++tm) {
                    const int kap = tm * kKterm + c * kP + p;
                    a[((size_t)(kap >> 3) * kChunkBytesA + k * 16) / 2 + (kap & 7)] = term[tm];
                }
            }
        }
    }
}

size_t mtd64_tc_matrix_bytes() { return tc::kABytes; }

cudaError_t launch_mtd64_tc(const float2* in, float* out, const void* a_mat, int in_ld, int out_ld, int cols, int n_slabs, int n_sms,
                            cudaStream_t st) {
    if (cols <= 0 || n_slabs <= 0) return cudaSuccess;
    const int tiles_per_slab = (cols + tc::kNr - 1) / tc::kNr;
    const long long n_items = (long long)tiles_per_slab * n_slabs;
    if (n_items > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    static size_t configured[64] = {};
    cudaError_t ce = ensure_dynamic_smem(mtd64_tc_kernel, (size_t)tc::kSmem, configured);
    if (ce != cudaSuccess) return ce;
    const int grid = (int)std::min<long long>(n_items, (long long)n_sms * 2);
    mtd64_tc_kernel<<<grid, tc::kThreads, tc::kSmem, st>>>(in, out, reinterpret_cast<const uint4*>(a_mat), in_ld, out_ld, cols, tiles_per_slab,
                                                           (int)n_items);
    return cudaGetLastError();
}

}  // namespace rb
