// cfar_kernels.cu -- K3: 2-D CA-CFAR (velocity axis dense, range axis on +-1 cells around velocity hits)
//                     with warp-ballot detection compaction.
//
// Replaces CW/executeCFAR.m:21-92, CW/Function_CFAR1D_sub.m:17-69 and
// CW/Function_CFAR1D_sub_fixCells.m:23-87.
//
//   stage V  (executeCFAR.m:28): every cell of the tested rows [n0+1, V-n0) is compared with
//            T_V * GO/SO(mean(left window), mean(right window)) along velocity; a window that falls
//            off the (cropped) axis is replaced by the other one (Function_CFAR1D_sub.m:30-39);
//            the compare is >= (:46).  Hits are compacted with __ballot_sync + one atomic per warp
//            into the detection list (kind = RB200_DET_V) and into a 1-bit-per-cell mask.
//   stage R  (executeCFAR.m:45-75): one thread per velocity hit (v,r) tests the range cells
//            {r-1,r,r+1} on row v and elects the first maximum among the passing ones (:64-73).
//            Because every write of the serial loop is "set to 1" the result is order independent;
//            the thread emits the elected cell (kind = RB200_DET_2D) unless a hit with a smaller
//            column elects the same cell (deterministic de-duplication through the bit mask).
//
// The same templates serve the float chain (row-major [v][r], list output) and the MATLAB-layout
// double entry points (column-major, dense 0/1 outputs, sums accumulated left-to-right so flags are
// bit-identical to the double-precision M-code).
#include "common.cuh"
#include "kernels.h"
#include "cfar_core.cuh"
#include "../../include/radar_b200.h"
#include <cstdlib>

namespace rb {

// Stage V.  One thread walks down one range column (rows [v_lo, v_hi), split over blockIdx.z into row
// segments): consecutive rows re-use 9 of the 10 reference rows from L1, every global access is a
// 128-byte row segment shared by the warp, and the 32 decisions of a warp for one row form one mask word
// (__ballot_sync).  ROWMAJOR: element (v,r) at v*R + r (float chain); otherwise column-major v + V*r.
template <typename T, bool ROWMAJOR>
__global__ void cfar_v_kernel(const T* __restrict__ rdm, const CfarParams p, T t_v, int rows_per_seg,
                              rb200_det* __restrict__ dets, int* __restrict__ det_count,
                              uint32_t* __restrict__ vmask, uint8_t* __restrict__ flagv, int* err_flag) {
    const int Rw = (p.R + 31) / 32;
    const int nv = p.v_hi - p.v_lo;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;      // blockDim.x is a multiple of 32
    const int lane = threadIdx.x & 31;
    const int word = r >> 5;
    int slo_, shi_;
    const bool in_range = r < p.R && cfar_seg_of(p.segs, r, p.R, &slo_, &shi_);     // columns outside every segment: no hits
    const T* slab_base = rdm + (size_t)slab * p.V * p.R;
    const T* col = ROWMAJOR ? slab_base + (size_t)p.v_lo * p.R + (in_range ? r : 0) : slab_base + p.v_lo + (size_t)p.V * (in_range ? r : 0);
    const ptrdiff_t sv = ROWMAJOR ? p.R : 1;
    const int y0 = blockIdx.z * rows_per_seg;
    const int y1 = min(nv, y0 + rows_per_seg);
    for (int vi = y0; vi < y1; ++vi) {
        const int v = p.v_lo + vi;
        bool hit = false;
        T amp = 0;
        if (in_range) {
            hit = cfar_decide<T>(col, sv, vi, nv, p.ref_v, p.guard_v, t_v, p.meth_v, err_flag);
            if (hit) amp = col[(ptrdiff_t)vi * sv];
        }
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        if (lane == 0 && word < Rw) vmask[((size_t)slab * p.V + v) * Rw + word] = ball;
        if (ball == 0) continue;
        if (flagv && hit) flagv[ROWMAJOR ? ((size_t)slab * p.V + v) * p.R + r : (size_t)slab * p.V * p.R + v + (size_t)p.V * r] = 1;
        int basei = 0;
        if (lane == 0) basei = atomicAdd(det_count, __popc(ball));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (hit) {
            const int slot = basei + __popc(ball & ((1u << lane) - 1u));
            if (slot < p.max_det) {
                rb200_det d;
                d.cpi = (uint32_t)(p.cpi0 + slab / p.n_lanes);
                d.r = (uint32_t)r;
                d.v = (uint16_t)v;
                d.lane = (uint8_t)(slab % p.n_lanes);
                d.kind = p.range_stage ? RB200_DET_V : (RB200_DET_V | RB200_DET_2D);   // executeCFAR.m:91
                d.amp = (float)amp;
                dets[slot] = d;
            }
        }
    }
}

// Stage V for the float chain with the reference rows staged through shared memory: a CTA owns 128 range
// columns and walks down the velocity axis in chunks of 32 rows; each chunk (plus ref+guard halo rows on both
// sides) is loaded once with 512-byte coalesced row segments, then every thread decides its 32 cells from
// shared memory.  HBM/L2 traffic drops from 11 loads per cell to (32+2H)/32.
#define RB_CFAR_TILE_ROWS 32
__global__ void __launch_bounds__(128)
cfar_v_tiled_kernel(const float* __restrict__ rdm, const CfarParams p, float t_v, int rows_per_seg,
                    rb200_det* __restrict__ dets, int* __restrict__ det_count, uint32_t* __restrict__ vmask, int* err_flag) {
    extern __shared__ float tile[];                 // [(32 + 2H)][128]
    const int H = p.ref_v + p.guard_v;
    const int Rw = (p.R + 31) / 32;
    const int nv = p.v_hi - p.v_lo;
    const int slab = blockIdx.y;
    const int r = blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int word = r >> 5;
    int slo_, shi_;
    const bool in_range = r < p.R && cfar_seg_of(p.segs, r, p.R, &slo_, &shi_);
    const float* col = rdm + ((size_t)slab * p.V + p.v_lo) * p.R + (in_range ? r : 0);
    const int y_begin = blockIdx.z * rows_per_seg;
    const int y_end = min(nv, y_begin + rows_per_seg);
    for (int y0 = y_begin; y0 < y_end; y0 += RB_CFAR_TILE_ROWS) {
        const int lo = max(0, y0 - H);
        const int hi = min(nv, y0 + RB_CFAR_TILE_ROWS + H);
        __syncthreads();                            // previous chunk fully consumed
        if (in_range)
            for (int y = lo; y < hi; ++y) tile[(y - lo) * 128 + threadIdx.x] = __ldg(col + (size_t)y * p.R);
        __syncthreads();
        const float* scol = tile + threadIdx.x - lo * 128;      // scol[y*128] == x[y][r] for y in [lo, hi)
        const int y1 = min(y_end, y0 + RB_CFAR_TILE_ROWS);
        for (int vi = y0; vi < y1; ++vi) {
            const int v = p.v_lo + vi;
            bool hit = false;
            float amp = 0.f;
            if (in_range) {
                hit = cfar_decide<float>(scol, 128, vi, nv, p.ref_v, p.guard_v, t_v, p.meth_v, err_flag);
                if (hit) amp = scol[vi * 128];
            }
            const unsigned ball = __ballot_sync(0xffffffffu, hit);
            if (lane == 0 && word < Rw) vmask[((size_t)slab * p.V + v) * Rw + word] = ball;
            if (ball == 0) continue;
            int basei = 0;
            if (lane == 0) basei = atomicAdd(det_count, __popc(ball));
            basei = __shfl_sync(0xffffffffu, basei, 0);
            if (hit) {
                const int slot = basei + __popc(ball & ((1u << lane) - 1u));
                if (slot < p.max_det) {
                    rb200_det d;
                    d.cpi = (uint32_t)(p.cpi0 + slab / p.n_lanes);
                    d.r = (uint32_t)r;
                    d.v = (uint16_t)v;
                    d.lane = (uint8_t)(slab % p.n_lanes);
                    d.kind = p.range_stage ? RB200_DET_V : (RB200_DET_V | RB200_DET_2D);   // executeCFAR.m:91
                    d.amp = amp;
                    dets[slot] = d;
                }
            }
        }
    }
}

template <typename T, bool ROWMAJOR>
__device__ __forceinline__ int cfar_elect(const T* __restrict__ row, ptrdiff_t sr, int r, int N, const CfarParams& p, T t_r, int* err_flag) {
    // first maximum among the passing cells of {r-1, r, r+1} (executeCFAR.m:50-73); -1 if none.  row points at the first
    // cell of the range segment, r is relative to it, N is the segment length.
    int best = -1;
    T bestv = 0;
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
        const int c = r + d;
        if (c < 0 || c >= N) continue;
        if (!cfar_decide<T>(row, sr, c, N, p.ref_r, p.guard_r, t_r, p.meth_r, err_flag)) continue;
        const T x = row[(ptrdiff_t)c * sr];
        if (best < 0 || x > bestv) { best = c; bestv = x; }
    }
    return best;
}

// Stage R: one thread per velocity hit.
template <typename T, bool ROWMAJOR>
__global__ void cfar_r_kernel(const T* __restrict__ rdm, const CfarParams p, T t_r,
                              const rb200_det* __restrict__ dets_v, const int* __restrict__ count_v,
                              rb200_det* __restrict__ dets_2d, int* __restrict__ count_2d,
                              const uint32_t* __restrict__ vmask, uint8_t* __restrict__ flag2d, int* err_flag) {
    // count_v[0] = hits so far, count_v[2] = first hit of this chunk
    const int i = count_v[2] + blockIdx.x * blockDim.x + threadIdx.x;
    int n = count_v[0];
    if (n > p.max_det) n = p.max_det;
    if (i >= n) return;
    const rb200_det h = dets_v[i];
    const int slab = (int)(h.cpi - p.cpi0) * p.n_lanes + h.lane;
    const int v = h.v, r = (int)h.r;
    const T* slab_base = rdm + (size_t)slab * p.V * p.R;
    const T* row = ROWMAJOR ? slab_base + (size_t)v * p.R : slab_base + v;
    const ptrdiff_t sr = ROWMAJOR ? 1 : p.V;
    int slo, shi;
    if (!cfar_seg_of(p.segs, r, p.R, &slo, &shi)) return;
    const T* srow = row + (ptrdiff_t)slo * sr;
    const int N = shi - slo;
    const int crel = cfar_elect<T, ROWMAJOR>(srow, sr, r - slo, N, p, t_r, err_flag);
    if (crel < 0) return;
    const int c = crel + slo;
    if (flag2d) {
        flag2d[ROWMAJOR ? ((size_t)slab * p.V + v) * p.R + c : (size_t)slab * p.V * p.R + v + (size_t)p.V * c] = 1;
        if (!dets_2d) return;
    }
    // de-duplicate: a velocity hit with a smaller column that elects the same cell owns the record
    const int Rw = (p.R + 31) / 32;
    const uint32_t* mrow = vmask + ((size_t)slab * p.V + v) * Rw;
    for (int rr = c - 1; rr < r; ++rr) {
        if (rr < slo) continue;
        if (!((mrow[rr >> 5] >> (rr & 31)) & 1u)) continue;
        if (cfar_elect<T, ROWMAJOR>(srow, sr, rr - slo, N, p, t_r, nullptr) == crel) return;
    }
    const int slot = atomicAdd(count_2d, 1);
    if (slot < p.max_det) {
        rb200_det d;
        d.cpi = h.cpi;
        d.r = (uint32_t)c;
        d.v = h.v;
        d.lane = h.lane;
        d.kind = RB200_DET_2D;
        d.amp = (float)row[(ptrdiff_t)c * sr];
        dets_2d[slot] = d;
    }
}

// 1-D CFAR over listed (row, col) cells of a column-major rows x cols matrix, detection along cols.
__global__ void cfar1d_kernel(const double* __restrict__ data, int rows, int cols, int ref, int guard, double T, int method,
                              const int* __restrict__ rows_fix, int n_rows_fix, const int* __restrict__ cols_fix, int n_cols_fix,
                              uint8_t* __restrict__ out, int* err_flag) {
    const int nr = rows_fix ? n_rows_fix : rows;
    const int nc = cols_fix ? n_cols_fix : cols;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nr * nc) return;
    const int ri = (int)(i % nr), ci = (int)(i / nr);
    const int row = rows_fix ? rows_fix[ri] - 1 : ri;
    const int y = cols_fix ? cols_fix[ci] - 1 : ci;
    if (cfar_decide<double>(data + row, rows, y, cols, ref, guard, T, method, err_flag)) out[row + (size_t)rows * y] = 1;
}

// ---------------------------------------------------------------------------------------------
template <typename T, bool ROWMAJOR>
static cudaError_t run_cfar(const T* rdm, const CfarParams& p, T t_r, T t_v, int n_slabs, rb200_det* dets_v, int* count_v,
                            rb200_det* dets_2d, int* count_2d, uint32_t* vmask, uint8_t* flag2d, uint8_t* flagv,
                            int* err_flag, cudaStream_t st) {
    const int Rw = (p.R + 31) / 32;
    const int nv = p.v_hi - p.v_lo;
    if (nv <= 0 || n_slabs <= 0) return cudaSuccess;
    (void)Rw;
    if (n_slabs > 65535) return cudaErrorInvalidConfiguration;
    // enough row segments to fill the machine when columns x slabs alone are few
    const int col_blocks = (p.R + 127) / 128;
    int segs = 1;
    while ((long long)col_blocks * n_slabs * segs < 1184 && nv / (segs * 2) >= 16 && segs < 64) segs *= 2;
    const int rows_per_seg = (nv + segs - 1) / segs;
    dim3 grid(col_blocks, n_slabs, (nv + rows_per_seg - 1) / rows_per_seg);
    const int H = p.ref_v + p.guard_v;
    static const bool no_tile = getenv("RB200_NO_CFAR_TILE") != nullptr;     // experiment switch, read once per process
    if (ROWMAJOR && sizeof(T) == 4 && !flagv && H <= 48 && !no_tile) {
        const size_t smem = (size_t)(RB_CFAR_TILE_ROWS + 2 * H) * 128 * sizeof(float);
        static size_t configured[64] = {};
        cudaError_t ce = ensure_dynamic_smem(cfar_v_tiled_kernel, smem, configured);
        if (ce != cudaSuccess) return ce;
        cfar_v_tiled_kernel<<<grid, 128, smem, st>>>(reinterpret_cast<const float*>(rdm), p, (float)t_v, rows_per_seg, dets_v, count_v, vmask, err_flag);
    } else {
        cfar_v_kernel<T, ROWMAJOR><<<grid, 128, 0, st>>>(rdm, p, t_v, rows_per_seg, dets_v, count_v, vmask, flagv, err_flag);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (p.range_stage) {
        cfar_r_kernel<T, ROWMAJOR><<<(p.max_det + 127) / 128, 128, 0, st>>>(rdm, p, t_r, dets_v, count_v, dets_2d, count_2d, vmask, flag2d, err_flag);
        e = cudaGetLastError();
    }
    return e;
}

cudaError_t launch_cfar_f32(const float* rdm, const CfarParams& p, float t_r, float t_v, int n_slabs, void* dets_v, int* count_v,
                            void* dets_2d, int* count_2d, uint32_t* vmask, uint8_t* flag2d, uint8_t* flagv, int* err_flag, cudaStream_t st) {
    return run_cfar<float, true>(rdm, p, t_r, t_v, n_slabs, (rb200_det*)dets_v, count_v, (rb200_det*)dets_2d, count_2d, vmask, flag2d, flagv, err_flag, st);
}

// range stage only (velocity hits already in dets_v / vmask, e.g. produced by the fused mtd_fast_kernel)
cudaError_t launch_cfar_r_f32(const float* rdm, const CfarParams& p, float t_r, void* dets_v, int* count_v, void* dets_2d, int* count_2d,
                              uint32_t* vmask, int* err_flag, cudaStream_t st) {
    if (!p.range_stage) return cudaSuccess;
    cfar_r_kernel<float, true><<<(p.max_det + 127) / 128, 128, 0, st>>>(rdm, p, t_r, (rb200_det*)dets_v, count_v, (rb200_det*)dets_2d, count_2d,
                                                                        vmask, nullptr, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_cfar_f64_colmajor(const double* rdm, const CfarParams& p, double t_r, double t_v, void* dets_v, int* count_v,
                                     uint32_t* vmask, uint8_t* flag2d, uint8_t* flagv, int* err_flag, cudaStream_t st) {
    return run_cfar<double, false>(rdm, p, t_r, t_v, 1, (rb200_det*)dets_v, count_v, nullptr, nullptr, vmask, flag2d, flagv, err_flag, st);
}

cudaError_t launch_cfar1d_f64(const double* data, int rows, int cols, int ref, int guard, double T, int method,
                              const int* rows_fix, int n_rows_fix, const int* cols_fix, int n_cols_fix,
                              uint8_t* out, int* err_flag, cudaStream_t st) {
    const size_t n = (size_t)(rows_fix ? n_rows_fix : rows) * (cols_fix ? n_cols_fix : cols);
    if (n == 0) return cudaSuccess;
    cfar1d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(data, rows, cols, ref, guard, T, method, rows_fix, n_rows_fix, cols_fix, n_cols_fix, out, err_flag);
    return cudaGetLastError();
}

}  // namespace rb
