// layout_kernels.cu -- MATLAB column-major split-complex double <-> device float layouts.
//
// The M functions exchange P x R matrices as column-major doubles (row = PRT index fastest,
// MP/fun_Process_MTD.m:6).  The device kernels want range-contiguous float lines.  These kernels do
// the precision change and the corner turn on the device through 32x33 shared-memory tiles so both
// sides are coalesced.
#include "common.cuh"
#include "kernels.h"

namespace rb {

// in: re/im column-major rows x cols (element (i,j) at i + rows*j) -> out row-major float2 [i][j]
// es: element stride of re / im in doubles (1 = split storage, 2 = interleaved complex with im = re + 1)
__global__ void z_to_planar_kernel(const double* __restrict__ re, const double* __restrict__ im, float2* __restrict__ out, int rows, int cols, int es) {
    __shared__ float2 tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int dj = threadIdx.y; dj < 32; dj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + dj;
        if (i < rows && j < cols) {
            const size_t k = ((size_t)i + (size_t)rows * j) * es;
            tile[dj][threadIdx.x] = make_float2((float)re[k], im ? (float)im[k] : 0.f);
        }
    }
    __syncthreads();
    for (int di = threadIdx.y; di < 32; di += blockDim.y) {
        const int i = i0 + di, j = j0 + threadIdx.x;
        if (i < rows && j < cols) out[(size_t)i * cols + j] = tile[threadIdx.x][di];
    }
}

__global__ void planar_to_z_kernel(const float2* __restrict__ in, double* __restrict__ re, double* __restrict__ im, int rows, int cols) {
    __shared__ float2 tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int di = threadIdx.y; di < 32; di += blockDim.y) {
        const int i = i0 + di, j = j0 + threadIdx.x;
        if (i < rows && j < cols) tile[di][threadIdx.x] = in[(size_t)i * cols + j];
    }
    __syncthreads();
    for (int dj = threadIdx.y; dj < 32; dj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + dj;
        if (i < rows && j < cols) {
            const size_t k = (size_t)i + (size_t)rows * j;
            const float2 v = tile[threadIdx.x][dj];
            re[k] = (double)v.x;
            im[k] = (double)v.y;
        }
    }
}

__global__ void f32_rowmajor_to_d_colmajor_kernel(const float* __restrict__ in, double* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int di = threadIdx.y; di < 32; di += blockDim.y) {
        const int i = i0 + di, j = j0 + threadIdx.x;
        if (i < rows && j < cols) tile[di][threadIdx.x] = in[(size_t)i * cols + j];
    }
    __syncthreads();
    for (int dj = threadIdx.y; dj < 32; dj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + dj;
        if (i < rows && j < cols) out[(size_t)i + (size_t)rows * j] = (double)tile[threadIdx.x][dj];
    }
}

// DMX sum / difference channel (CW/DMX_SignalProcessing_main_xzr.m:421-426,462-465): magnitudes of the two beams,
// row-major float [beam][rows][ld] (columns c0 .. c0+cols-1 used) -> column-major double sum = |L|+|R| with rows
// [0, n_blank] and [rows-n_blank, rows) zeroed, diff = |R|-|L| (not blanked).
__global__ void dmx_combine_kernel(const float* __restrict__ left, const float* __restrict__ right, int ld, int c0, double* __restrict__ sum,
                                   double* __restrict__ diff, int rows, int cols, int n_blank) {
    __shared__ float tl[32][33], tr[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int di = threadIdx.y; di < 32; di += blockDim.y) {
        const int i = i0 + di, j = j0 + threadIdx.x;
        if (i < rows && j < cols) {
            tl[di][threadIdx.x] = left[(size_t)i * ld + c0 + j];
            tr[di][threadIdx.x] = right[(size_t)i * ld + c0 + j];
        }
    }
    __syncthreads();
    for (int dj = threadIdx.y; dj < 32; dj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + dj;
        if (i < rows && j < cols) {
            const double a = (double)tl[threadIdx.x][dj], b = (double)tr[threadIdx.x][dj];
            const bool blank = n_blank >= 0 && (i <= n_blank || i >= rows - n_blank);
            if (sum) sum[(size_t)i + (size_t)rows * j] = blank ? 0.0 : a + b;
            if (diff) diff[(size_t)i + (size_t)rows * j] = b - a;
        }
    }
}

__global__ void u8_to_d_kernel(const uint8_t* __restrict__ in, double* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] ? 1.0 : 0.0;
}

// out = in with rows [lo,hi] (0-based inclusive) zeroed; column-major double (fun_0v_pressing on its own)
__global__ void zero_rows_kernel(const double* __restrict__ in, double* __restrict__ out, int rows, int cols, int lo, int hi) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    const int r = (int)(i % rows);
    out[i] = (r >= lo && r <= hi) ? 0.0 : in[i];
}

static dim3 tgrid(int rows, int cols) { return dim3((rows + 31) / 32, (cols + 31) / 32, 1); }

cudaError_t launch_z_to_planar(const double* re, const double* im, float2* out, int rows, int cols, cudaStream_t st, int elem_stride) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    z_to_planar_kernel<<<tgrid(rows, cols), dim3(32, 8), 0, st>>>(re, im, out, rows, cols, elem_stride);
    return cudaGetLastError();
}
cudaError_t launch_planar_to_z(const float2* in, double* re, double* im, int rows, int cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    planar_to_z_kernel<<<tgrid(rows, cols), dim3(32, 8), 0, st>>>(in, re, im, rows, cols);
    return cudaGetLastError();
}
cudaError_t launch_f32_rowmajor_to_d_colmajor(const float* in, double* out, int rows, int cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    f32_rowmajor_to_d_colmajor_kernel<<<tgrid(rows, cols), dim3(32, 8), 0, st>>>(in, out, rows, cols);
    return cudaGetLastError();
}
cudaError_t launch_dmx_combine(const float* left, const float* right, int ld, int c0, double* sum, double* diff, int rows, int cols,
                               int n_blank, cudaStream_t st) {
    if (rows <= 0 || cols <= 0 || (!sum && !diff)) return cudaSuccess;
    dmx_combine_kernel<<<tgrid(rows, cols), dim3(32, 8), 0, st>>>(left, right, ld, c0, sum, diff, rows, cols, n_blank);
    return cudaGetLastError();
}
cudaError_t launch_u8_to_d(const uint8_t* in, double* out, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    u8_to_d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
    return cudaGetLastError();
}
cudaError_t launch_zero_rows_d_colmajor(const double* in, double* out, int rows, int cols, int lo, int hi, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    const size_t n = (size_t)rows * cols;
    zero_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, rows, cols, lo, hi);
    return cudaGetLastError();
}

// |in[row][col]| (complex planar, row-major) -> out (rows x cols, column-major double)
__global__ void abs_planar_to_d_colmajor_kernel(const float2* __restrict__ in, double* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        float v = 0.f;
        if (r < rows && c < cols) {
            const float2 x = in[(size_t)r * cols + c];
            v = hypotf(x.x, x.y);
        }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = (double)tile[threadIdx.x][j];
    }
}
cudaError_t launch_abs_planar_to_d_colmajor(const float2* in, double* out, int rows, int cols, cudaStream_t st) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, 1), block(32, 8, 1);
    abs_planar_to_d_colmajor_kernel<<<grid, block, 0, st>>>(in, out, rows, cols);
    return cudaGetLastError();
}

}  // namespace rb
