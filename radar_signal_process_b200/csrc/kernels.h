// kernels.h -- host-callable launchers of the CUDA kernels (internal to libradar_b200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "common.cuh"

namespace rb {

// ---- K1 pulse compression (pc_kernels.cu)
int pc_tile_lanes(int nt, bool wire);
void pc_build_twiddles(int nt, std::vector<float2>& tw);   // stage-major twiddle table for tile size nt     // lines per CTA for FFT tile size nt (0 = unsupported)
cudaError_t launch_pc_fft(int nt, bool wire, const PcParams& p, int n_tiles, int n_groups, cudaStream_t st);
// persistent TMA-prefetch variant: wire format, 16 channels, 256-sample tiles
// h_entries: float2 entries of PcParams::hperm to keep resident in shared memory (all segment spectra)
cudaError_t launch_pc_fft_tma(const PcParams& p, int n_tiles, int n_groups, int n_sms, int ctas_per_sm, int h_entries, cudaStream_t st);
// warp-private variant (pcw_kernel.cu): one CTA per SM, 16 warps, two lines per thread, tensor-map TMA with 128-byte swizzle
bool pcw_plan_supported(const PcParams& p, int n_segs, int h_entries);
cudaError_t launch_pcw(const PcParams& p, int n_tiles, int n_groups, int n_sms, int h_entries, bool shared_sm, cudaStream_t st);
cudaError_t launch_pc_direct(bool wire, const PcParams& p, const float2* taps, int seg_idx, int out_len, int n_lines, cudaStream_t st);
cudaError_t launch_pc_zero_cols(float2* out, size_t n_lines, int R, int c0, int c1, cudaStream_t st);
cudaError_t launch_unpack(const int16_t* raw, float2* out, int n_groups, int P, int R, int C, cudaStream_t st);

// ---- DBF weighting fused with the unpack (dbf_kernel.cu): wire int16 -> float2 planar [cpi][beam][prt][range]
cudaError_t launch_dbf(const int16_t* raw, float2* out, const float2* W, int n_beams, int n_ch, int n_groups, int P, int R, cudaStream_t st);

cudaError_t launch_unpack_dbf24(const uint8_t* bytes, float2* out, int n_prt, int n, int ncol, int row_bytes, size_t prt_bytes, cudaStream_t st);

// ---- K2 MTD (mtd_kernels.cu)
bool mtd_has_fast_path(int P);
bool mtd_fast_fuses_cfar(int P);          // the P = R*R shared-memory kernels can run the velocity CFAR on their tile
cudaError_t launch_mtd(const MtdParams& p, int n_slabs, cudaStream_t st);
int mtd_generic_max_p();

// ---- K2 for P = 64 fused with the velocity CFAR stage, and its sparse range stage (mtd64_kernel.cu)
bool mtd64_fused_supported(int P, int ref_v, int guard_v, int n0, int mti_lag);
cudaError_t launch_mtd64(const Mtd64Params& p, int n_slabs, bool with_cfar, cudaStream_t st);
cudaError_t launch_mtd64_tma(const Mtd64Params& p, int n_slabs, int n_sms, int ctas_per_sm, cudaStream_t st);   // persistent, TMA-staged tiles (even in_ld / cols)
cudaError_t launch_wait_flag(const int* flag, int* err_flag, cudaStream_t st);     // RB200_SPLIT: stream waits until *flag != 0
cudaError_t launch_cfar_r64(const float* rdm, const CfarParams& p, float t_r, const void* slot_v, int* slot_count, void* dets_v,
                            void* dets_2d, int* gcount, const unsigned long long* colmask, int cols_ld, int* err_flag, int n_blocks,
                            cudaStream_t st);

// ---- experiment: P = 64 slow-time DFT as a tcgen05 GEMM (mtd64_tc_kernel.cu), RDM only
void mtd64_tc_build_matrix(const float* window, int zv_lo, int zv_hi, std::vector<uint16_t>& a);
size_t mtd64_tc_matrix_bytes();
cudaError_t launch_mtd64_tc(const float2* in, float* out, const void* a_mat, int in_ld, int out_ld, int cols, int n_slabs, int n_sms,
                            cudaStream_t st);

// ---- fused persistent PC + MTD64 + velocity CFAR for a whole batch (chain64_kernel.cu)
cudaError_t launch_chain64(const Chain64Params& q, int n_sms, cudaStream_t st);

// ---- single-pass chain for P = 64, 16 lanes: unpack + PC + MTD + 0-v + velocity CFAR with the PC intermediate in shared memory
// (onepass_kernel.cu)
int onepass_tile_valid(int n_taps);                 // V: valid lags per 256-sample tile (multiple of 4, <= 192)
int onepass_planar_pitch(int R, int V);             // 8-byte elements per (lane, PRT pair) row of the lane planes
size_t onepass_planar_bytes(int n_cpi, int R, int V);
cudaError_t onepass_planar_init(void* planar, size_t bytes, cudaStream_t st);     // pad columns = offset-binary zero (once per allocation)
cudaError_t launch_deinterleave(const void* raw, void* planar, int n_cpi, int R, int V, cudaStream_t st);   // wire -> lane planes
cudaError_t launch_onepass(const OnePassParams& p, void* planar, int n_sms, cudaStream_t st);

// ---- K3 CFAR (cfar_kernels.cu)
// chain variant: float RDM [slab][V][R] row-major -> velocity-hit list + 2-D list (+ optional dense uint8 flags)
cudaError_t launch_cfar_f32(const float* rdm, const CfarParams& p, float t_r, float t_v, int n_slabs, void* dets_v, int* count_v,
                            void* dets_2d, int* count_2d, uint32_t* vmask, uint8_t* flag2d, uint8_t* flagv, int* err_flag, cudaStream_t st);
cudaError_t launch_cfar_r_f32(const float* rdm, const CfarParams& p, float t_r, void* dets_v, int* count_v, void* dets_2d, int* count_2d,
                              uint32_t* vmask, int* err_flag, cudaStream_t st);
// MATLAB variant: one double V x R column-major matrix -> dense uint8 flags (column-major)
cudaError_t launch_cfar_f64_colmajor(const double* rdm, const CfarParams& p, double t_r, double t_v, void* dets_v, int* count_v,
                                     uint32_t* vmask, uint8_t* flag2d, uint8_t* flagv, int* err_flag, cudaStream_t st);
// 1-D CFAR along the second dimension of a column-major rows x cols double matrix, all or listed (1-based) cells
cudaError_t launch_cfar1d_f64(const double* data, int rows, int cols, int ref, int guard, double T, int method,
                              const int* rows_fix, int n_rows_fix, const int* cols_fix, int n_cols_fix,
                              uint8_t* out, int* err_flag, cudaStream_t st);

// ---- post-CFAR measurement (measure_kernels.cu), all arrays column-major double on the device
cudaError_t launch_flag_compaction(const double* flags, int V, int R, int* col_start, int* total, cudaStream_t st);
cudaError_t launch_measure(const double* sum, const double* diff, const double* flags, const double* rScale, const double* vScale,
                           const double* kValues, int kRows, int V, int R, int extra, int rTimes, int vTimes, int n0, double deltaR,
                           double deltaV, double beamPosNum, double beamAngleStep, int freInd, double eleComp, double eleSysErr,
                           const int* col_start, double* out_r, double* out_v, double* out_e, int* err_flag, cudaStream_t st);

// ---- layout conversion (layout_kernels.cu): MATLAB column-major split double <-> device layouts
cudaError_t launch_z_to_planar(const double* re, const double* im, float2* out, int rows, int cols, cudaStream_t st, int elem_stride = 1);          // out[row][col]
cudaError_t launch_planar_to_z(const float2* in, double* re, double* im, int rows, int cols, cudaStream_t st);                // in[row][col]
cudaError_t launch_abs_planar_to_d_colmajor(const float2* in, double* out, int rows, int cols, cudaStream_t st);   // |in[row][col]|
cudaError_t launch_f32_rowmajor_to_d_colmajor(const float* in, double* out, int rows, int cols, cudaStream_t st);
cudaError_t launch_dmx_combine(const float* left, const float* right, int ld, int c0, double* sum, double* diff, int rows, int cols,
                               int n_blank, cudaStream_t st);
cudaError_t launch_u8_to_d(const uint8_t* in, double* out, size_t n, cudaStream_t st);
cudaError_t launch_zero_rows_d_colmajor(const double* in, double* out, int rows, int cols, int lo, int hi, cudaStream_t st);

}  // namespace rb
