/*
 * radar_b200.h -- C ABI of libradar_b200.so: the B200-native (sm_100a) per-frame detection chain
 *
 *     int16 DDC unpack -> pulse compression -> MTD (Kaiser, slow-time FFT, fftshift, |.|)
 *                      -> zero-velocity suppression -> 2-D CA-CFAR -> detection list
 *
 * Drop-in boundary for the MATLAB functions of XuZerui2023/Radar-Signal-Process.  Each entry point
 * names the reference interface it replaces (file:line, relative to the reference root;
 * MP = MatlabProcess_xuzerui, CW = MatlabProcess_xuzerui/CFAR_WangCai).  The MEX gateways in mex/
 * bind exactly these symbols; INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions cross the boundary.
 *   - every call returns an rb200_status (0 = OK); rb200_last_error() gives the text.
 *   - "_z" entry points take MATLAB split-complex double, column-major (row index fastest);
 *     "_d" entry points take real double column-major.  Host pointers only.  Conversion to the
 *     device's float32 layout happens on the device, not on the host.
 *   - a context is bound to one CUDA device and one stream; it is not thread-safe; distinct
 *     contexts are independent (one per rank / per MATLAB session).
 *   - the caller owns every buffer it passes; the library owns only context-internal scratch.
 *   - there is no CPU fallback: without a CUDA device rb200_create() fails with RB200_ERR_CUDA.
 */
#ifndef RADAR_B200_H
#define RADAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_VERSION_MAJOR 0
#define RB200_VERSION_MINOR 1

typedef struct rb200_ctx rb200_ctx;

typedef enum {
    RB200_OK = 0,
    RB200_ERR_ARG = 1,          /* bad argument (null pointer, negative size, unknown enum)           */
    RB200_ERR_CUDA = 2,         /* CUDA runtime error (sticky; text in rb200_last_error)              */
    RB200_ERR_INDEX = 3,        /* the M-code would raise "Index exceeds array bounds"                */
    RB200_ERR_DIM_MISMATCH = 4, /* the M-code would raise a dimension-mismatch error                  */
    RB200_ERR_NO_WAVEFORM = 5,  /* rb200_set_waveform() has not been called                           */
    RB200_ERR_UNSUPPORTED = 6,  /* outside the implemented envelope (documented per call)             */
    RB200_ERR_OVERFLOW = 7      /* detection list truncated: *n_det holds the true count              */
} rb200_status;

/* ---- chain geometry and detector parameters --------------------------------------------------
 * Mirrors the literals the reference scatters through its scripts: geometry
 * (MP/fun_MTD_produce.m:8,41-44), Kaiser beta (MP/fun_Process_MTD.m:13), zero-velocity divisor
 * (MP/fun_0v_pressing.m:4-6 -> 150, CW/fun_0v_pressing.m:5 -> 20), MTI lag (MP/fun_Process_MTI.m:20-21),
 * CFAR parameters (CW/main_cfar.m:40-58; argument order of CW/executeCFAR.m:1-2).                 */
typedef struct {
    int32_t struct_size;     /* = sizeof(rb200_config), for ABI evolution                            */
    int32_t n_prt;           /* P: PRTs per CPI (slow time)                                          */
    int32_t n_range;         /* R: range cells per PRT (fast time)                                   */
    int32_t n_lanes;         /* C: channels/beams interleaved in the wire format                     */
    int32_t max_cpi;         /* largest n_cpi a single rb200_chain_i16 call will pass                */
    int32_t max_det;         /* capacity of the device detection list per call                       */
    int32_t zero_v_div;      /* 150 (MP/, MTD/), 20 (CW/), 0 = no zero-velocity suppression          */
    int32_t mti_lag;         /* 0 = off; 30 = fun_Process_MTI, applied after PC, before the window   */
    double  kaiser_beta;     /* 8                                                                    */
    int32_t cfar_ref_r, cfar_guard_r, cfar_method_r;   /* range axis: ref cells, guard cells, 0=GO 1=SO */
    int32_t cfar_ref_v, cfar_guard_v, cfar_method_v;   /* velocity axis                              */
    double  cfar_t_r, cfar_t_v;                        /* threshold factors                          */
    int32_t cfar_n0;         /* MTD_0_num: rows 1..n0+1 and the last n0 are not tested               */
    int32_t cfar_range_stage;/* rCFARDetect_Flag                                                     */
    int32_t chunk_cpi;       /* CPIs per PC->MTD->CFAR pass (L2 residency); 0 = library default      */
    int32_t reserved;
} rb200_config;

/* ---- waveform plan ----------------------------------------------------------------------------
 * One entry per waveform segment of a PRT (MP/fun_lss_pulse_compression.m:6-8,14-16,31,36-37;
 * MTD/fun_lss_pulse_compression.m:23-25,47-51,58-65).                                            */
typedef enum {
    RB200_SEG_MF = 0,   /* fun_pulse_compression(taps, x) then (L : L+out_len-1)                     */
    RB200_SEG_FIR = 1,  /* filter(taps,1,x) * scale                                                  */
    RB200_SEG_MF_CIRC = 2 /* circular matched filter ifft(fft(x,N).*conj(fft(taps,N))), N = out_len in {256,512,4096},
                           * in_len <= N zero-padded (CW/DMX_SignalProcessing_main_xzr.m:202,348-353)      */
} rb200_seg_kind;

typedef enum {
    RB200_ALIGN_LEADING_EDGE = 0, /* MF: output n <-> lag n (peak at the echo's first sample)        */
    RB200_ALIGN_DELAYED = 1,      /* FIR, MP/ 5-arg rule: causal output left delayed (:31)           */
    RB200_ALIGN_GRPDELAY = 2      /* FIR, MTD/ 9-arg rule: circshift(y,-round(mean(grpdelay))) (:47-51) */
} rb200_seg_align;

typedef struct {
    int32_t in_start, in_len;    /* 0-based first column and length of the input slice               */
    int32_t out_start, out_len;  /* where the result lands in the output PRT (out_len <= in_len)     */
    int32_t kind;                /* rb200_seg_kind                                                   */
    int32_t align;               /* rb200_seg_align                                                  */
    int32_t n_taps;
    int32_t reserved;
    const double* taps_re;       /* MF: reference pulse s0 (un-flipped, un-conjugated); FIR: b       */
    const double* taps_im;       /* may be NULL (real taps)                                          */
    double scale;                /* multiplies the segment output (FIR: 1/1.2; MF: 1)                */
} rb200_segment;

/* One detection.  kind bit0 = velocity-stage hit (cfarResultFlag_MatrixV), bit1 = final 2-D flag
 * (cfarResultFlag_Matrix).  A cell may appear in two records (one per bit); with the range stage off
 * every velocity hit carries both bits (CW/executeCFAR.m:91).  v and r are 0-based.               */
typedef struct {
    uint32_t cpi;
    uint32_t r;
    uint16_t v;
    uint8_t  lane;
    uint8_t  kind;
    float    amp;
} rb200_det;

#define RB200_DET_V    1u
#define RB200_DET_2D   2u

/* ---- lifetime ---------------------------------------------------------------------------------*/
int rb200_version(void);                                      /* major*100 + minor                  */
int rb200_create(rb200_ctx** out, int device, const rb200_config* cfg);
int rb200_destroy(rb200_ctx* ctx);
const char* rb200_last_error(const rb200_ctx* ctx);           /* NULL ctx -> last create() error    */
int rb200_get_config(const rb200_ctx* ctx, rb200_config* out);

/* One process-wide, reference-counted context per device (default configuration) for hosts that load several thin
 * gateways into one process: every MEX gateway of mex/ (one binary per M-function name, e.g. executeCFAR of
 * CW/main_cfar.m:90 and fun_MTD_produce of MP/main_produce_dataset_win_xzr.m:37 in one MATLAB session) acquires THIS
 * context instead of creating a private one, so a session holds one CUDA context, one set of device scratch and one plan
 * cache.  acquire: creates on first use, otherwise bumps the count; release: destroys the context at count 0.
 * Same single-thread rule as any context (the interpreter calls mexFunction on its main thread).                      */
int rb200_shared_context_acquire(rb200_ctx** out, int device);
int rb200_shared_context_release(int device);
/* Opaque 64-bit tag naming the waveform plan resident in a context (0 = unknown).  rb200_set_waveform clears it; a caller
 * that built the plan stores its own hash and can later tell -- across gateways sharing the context -- whether the plan
 * it needs is still the resident one.                                                                                  */
int rb200_set_plan_tag(rb200_ctx* ctx, uint64_t tag);
uint64_t rb200_get_plan_tag(const rb200_ctx* ctx);
int rb200_set_cfar(rb200_ctx* ctx, const rb200_config* cfg);

/* Range segments of the chain's CFAR = the local function fun_CFARflag of CW/main_cfar.m:142-161 (columns 1:82, 83:318,
 * 319:868 there): executeCFAR runs on every segment separately, so range windows never straddle a waveform segment, and
 * columns outside every segment are left 0.  lo/hi: 0-based half-open column ranges, ascending and disjoint, n <= 4;
 * n = 0 restores one segment covering the whole PRT.  A segment too short for the range windows raises
 * RB200_ERR_INDEX at the next chain call, as Function_CFAR1D_sub_fixCells.m:39-58 would.                             */
int rb200_set_cfar_segments(rb200_ctx* ctx, const int32_t* lo, const int32_t* hi, int n);  /* update only the cfar_* fields      */

/* Precompute and keep resident the reference spectra of every segment.  Replaces the per-PRT
 * fft(h,n) of MP/fun_pulse_compression.m:20.                                                      */
int rb200_set_waveform(rb200_ctx* ctx, const rb200_segment* segs, int nseg);

/* iSTC gain curve in dB, zero-extended to n_range; n = 0 disables.  MP/fun_iSTC.m:6-9,14.         */
int rb200_set_stc(rb200_ctx* ctx, const double* stc_db, int n);

/* ---- MATLAB-layout entry points (one per reference function) -----------------------------------*/

/* signal_PC = fun_pulse_compression(s0, s_echo)        MP/fun_pulse_compression.m:1
 * out_* hold L+M-1 samples.                                                                       */
int rb200_pulse_compression_z(rb200_ctx* ctx, const double* s0_re, const double* s0_im, int L,
                              const double* echo_re, const double* echo_im, int M,
                              double* out_re, double* out_im);

/* s_PC_0 = fun_lss_pulse_compression(echo, ...)        MP/fun_lss_pulse_compression.m:3,
 *                                                      MTD/fun_lss_pulse_compression.m:17
 * Uses the plan given to rb200_set_waveform.  echo and out are P x R column-major.                */
int rb200_lss_pulse_compression_z(rb200_ctx* ctx, const double* echo_re, const double* echo_im,
                                  int P, int R, double* out_re, double* out_im);

/* MTD_Signal = fun_Process_MTD(ProSignal, Len_PRT, Num_PRTperFrame)   MP/fun_Process_MTD.m:3
 * ProSignal is rows x cols; out is Num_PRTperFrame x Len_PRT real.                                */
int rb200_process_mtd_z(rb200_ctx* ctx, const double* re, const double* im, int rows, int cols,
                        int len_prt, int num_prt, double beta, double* out);

/* MTD = fun_0v_pressing(MTD)                           MP/fun_0v_pressing.m:2 (div 150),
 *                                                      CW/fun_0v_pressing.m:2 (div 20)            */
int rb200_zero_v_pressing_d(rb200_ctx* ctx, const double* mtd, int P, int R, int div, double* out);

/* MTD_Signal = fun_MTD_produce(echo[,params])          MP/fun_MTD_produce.m:3, MTD/fun_MTD_produce.m:12
 * PC (current plan) -> [MTI] -> MTD -> 0-v on a P x R column-major frame; out is P x R real.      */
int rb200_mtd_produce_z(rb200_ctx* ctx, const double* echo_re, const double* echo_im, int P, int R,
                        double beta, int zero_v_div, double* out);

/* MTD_crop = fun_MTD_produce(echo)(row_lo:row_hi, :)    MP/main_produce_dataset_win_xzr.m:37-40 (rows 691:845 of 1536)
 * Same result as rb200_mtd_produce_z followed by the caller's row crop (1-based, inclusive), computed with the slow-time
 * transform first and the pulse compression on the kept rows only; out is (row_hi-row_lo+1) x R real, column-major.
 * Errors: RB200_ERR_INDEX when the crop leaves 1..P (MATLAB: "Index in position 1 exceeds array bounds").          */
int rb200_mtd_produce_rows_z(rb200_ctx* ctx, const double* echo_re, const double* echo_im, int P, int R,
                             double beta, int zero_v_div, int row_lo, int row_hi, double* out);

/* The same two entry points for INTERLEAVED complex input (MATLAB -R2018a mxGetComplexDoubles, Octave's and numpy's native
 * storage): echo_ri holds P x R column-major (re, im) pairs; one H2D copy, no host-side split.                     */
int rb200_mtd_produce_c(rb200_ctx* ctx, const double* echo_ri, int P, int R, double beta, int zero_v_div, double* out);
int rb200_mtd_produce_rows_c(rb200_ctx* ctx, const double* echo_ri, int P, int R, double beta, int zero_v_div,
                             int row_lo, int row_hi, double* out);

/* F = Function_CFAR1D_sub(data, ref, save, T, method)  CW/Function_CFAR1D_sub.m:1
 * data/out are rows x cols column-major; detection runs along columns index (second dim).         */
int rb200_cfar1d_sub_d(rb200_ctx* ctx, const double* data, int rows, int cols, int ref, int guard,
                       double T, int method, double* out);

/* F = Function_CFAR1D_sub_fixCells(data, ref, save, T, method, rowCellsFix, colCellsFix)
 *                                                      CW/Function_CFAR1D_sub_fixCells.m:1
 * rows_fix / cols_fix are 1-based like the M arguments.                                           */
int rb200_cfar1d_fix_d(rb200_ctx* ctx, const double* data, int rows, int cols, int ref, int guard,
                       double T, int method, const int32_t* rows_fix, int n_rows_fix,
                       const int32_t* cols_fix, int n_cols_fix, double* out);

/* [F, FV] = executeCFAR(mtd, refR, saveR, T_R, methR, refV, saveV, T_V, methV, n0, rFlag)
 *                                                      CW/executeCFAR.m:1
 * mtd, out_flag, out_flag_v are V x R column-major; out_flag_v may be NULL.                       */
int rb200_execute_cfar_d(rb200_ctx* ctx, const double* mtd, int V, int R,
                         int ref_r, int guard_r, double t_r, int method_r,
                         int ref_v, int guard_v, double t_v, int method_v,
                         int n0, int range_stage, double* out_flag, double* out_flag_v);

/* ---- batched wire-format entry points (benchmark / multi-GPU path) ------------------------------
 * raw: int16 [cpi][prt][range][lane][I,Q]  (FrameDataRead_xzr.m:150-156 per PRT, payload only).   */

/* Parity/debug: unpacked samples as float2 [cpi][lane][prt][range] (re,im interleaved). Host ptrs. */
int rb200_unpack_ddc_i16(rb200_ctx* ctx, const int16_t* raw, int n_cpi, float* out_ri);

/* The whole chain for n_cpi CPIs.  raw / rdm_out / dets may each be host or device pointers
 * (detected with cudaPointerGetAttributes); rdm_out ([cpi][lane][v][range] float) and dets may be
 * NULL.  *n_det (host int) receives the number of detections found; at most max_det are stored
 * (RB200_ERR_OVERFLOW if more).  stream: a cudaStream_t cast to void*, NULL = the context stream.
 * The call returns after the work has completed on the stream.                                    */
int rb200_chain_i16(rb200_ctx* ctx, const int16_t* raw, int n_cpi, float* rdm_out,
                    rb200_det* dets, int* n_det, void* stream);

/* Same as rb200_chain_i16 with device-resident raw/rdm_out, but only enqueues the kernels (no
 * synchronisation, no D2H): used to time the device chain with caller-owned events.  The
 * detection count/list stay in context memory until rb200_chain_fetch.                            */
int rb200_chain_enqueue(rb200_ctx* ctx, const int16_t* raw_dev, int n_cpi, float* rdm_dev, void* stream);
int rb200_chain_fetch(rb200_ctx* ctx, rb200_det* dets_host, int* n_det);
/* Zero-copy view of the detection lists of the last chain call, in DEVICE memory (valid until the next chain call on this
 * context): 2-D detections (the flags of CW/executeCFAR.m:78-89) and velocity-stage hits (:28-31), with their counts
 * (waits for the chain to finish).  For consumers that stay on the device -- the multi-GPU gather hands these pointers to
 * NCCL directly, no host bounce.  Counts are clipped to rb200_config.max_det.                              */
int rb200_chain_dets_device(rb200_ctx* ctx, const rb200_det** dets_2d, int* n_2d, const rb200_det** dets_v, int* n_v);

/* Device-side intermediates of the last chain call (parity tests): pulse-compressed samples as
 * float2 [cpi][lane][prt][range] for the last processed chunk.  The default chain for 64 PRT x 16 int16 lanes keeps that
 * intermediate in shared memory (single-pass kernel: MP/fun_MTD_produce.m:67-79 in one pass); rb200_set_debug_keep_pc(ctx, 1)
 * selects the pipeline that materialises it in device memory so that it can be fetched.                                 */
int rb200_set_debug_keep_pc(rb200_ctx* ctx, int on);
int rb200_debug_fetch_pc(rb200_ctx* ctx, int cpi_in_chunk, float* out_ri);

/* Milliseconds between the first and last kernel of the last chain call (CUDA events).            */
int rb200_last_device_ms(const rb200_ctx* ctx, float* ms);
/* Number of kernels the last call launched (bench.py's gpu_launches).                             */
int rb200_last_launch_count(const rb200_ctx* ctx, int* n);

/* Per-stage device timing for the roofline report: when on, every chunk of a chain call is bracketed
 * by CUDA events on the launching stream (before PC, after PC, after MTD, after CFAR).
 * rb200_get_stage_ms sums the elapsed milliseconds of {PC, MTD, CFAR} over all chunks recorded since
 * the last call and reports how many chunks / CPIs they covered.                                    */
int rb200_set_stage_timing(rb200_ctx* ctx, int on);
int rb200_get_stage_ms(rb200_ctx* ctx, float ms[3], int* n_chunks, int* n_cpis);

/* ---- "next" rows ---------------------------------------------------------*/

/* f1: DBF weighting fused into the unpack: beams = sig_C * W.' (FrameDataRead_xzr.m:158, non-conjugate
 * transpose).  w_re/w_im: n_beams x n_lanes column-major (DBF_coeffs_data_C, bin_to_mat_xzr.m:23-29);
 * w_im may be NULL; n_beams = 0 disables.  When enabled, every lane index downstream of the unpack
 * (rdm_out [cpi][beam][v][range], rb200_det.lane) is a beam index and buffers are sized with n_beams.   */
int rb200_set_dbf(rb200_ctx* ctx, const double* w_re, const double* w_im, int n_beams);

/* f2: capture-file byte stream + per-PRT frame parser (host code, no GPU needed).
 * Replaces read_continuous_file_stream.m:22 (files <dir>/1.00000N.bin, DataFullPathGen.m:10-16, one logical stream
 * with the reference's file-index behaviour) and the DDC branch of FrameDataRead_xzr.m:57-198.
 * rb200_reader_next_frame_ddc parses up to n_prt PRTs (64 B head | 128 B realtime | payload padded to 64 B | 64 B
 * tail) and stores their int16 payloads as [prt][range][channel][I,Q] in raw_out (typically pinned memory handed
 * straight to rb200_chain_i16).  *prts_read < n_prt with *end_of_stream = 1 when the stream ends or a PRT is
 * inconsistent (the M-code's "frame not completed").  Optional per-PRT header fields: frame number (head word 1),
 * servo angle (low 16 bits of word 5, 0.1 deg), 64-bit timer (words 9-10).                                        */
typedef struct rb200_reader rb200_reader;
int rb200_reader_open(rb200_reader** out, const char* dir);
int rb200_reader_close(rb200_reader* r);
const char* rb200_reader_last_error(const rb200_reader* r);
int rb200_reader_state(const rb200_reader* r, int* file_index, long long* pos);
int rb200_reader_next_frame_ddc(rb200_reader* r, int n_prt, int n_range, int n_channels, int16_t* raw_out,
                                uint32_t* frame_no, uint16_t* servo_angle, uint64_t* timer_cnt,
                                int* prts_read, int* end_of_stream);
/* Same for DBF-type frames (data_type 2, FrameDataRead_xzr.m:111-119): payload_out receives n_prt padded PRT payloads
 * (n_range * (6*n_channels + pad) bytes, rounded up to 64) back to back -- the input layout of rb200_chain_dbf24.   */
int rb200_reader_next_frame_dbf24(rb200_reader* reader, int n_prt, int n_range, int n_channels, uint8_t* payload_out,
                                  uint32_t* frame_no, uint16_t* servo_angle, uint64_t* timer_cnt,
                                  int* prts_read, int* end_of_stream);

/* f3: [rEst, vEst, eleEst] = motionParaMeasure(sum, diff, flags, extraDots, rScale, deltaR, rInterpTimes, vScale, deltaV,
 *          vInterpTimes, kValues, beamPosNum, beamAngleStep, freInd, eleAngleComp, eleAngleSysErr, MTD_0_num)
 * Replaces CW/motionParaMeasure.m:1.  sum / diff / flags: V x R column-major double; r_scale: R values; v_scale: V
 * values; k_values: k_rows x k_cols column-major (angle_KvalueGen.m).  Outputs hold one value per flagged cell in
 * MATLAB find() order; capacity = room in each output array; *n_out = number of flagged cells (RB200_ERR_OVERFLOW
 * if larger than capacity).  extra_dots <= 16.                                                                  */
int rb200_motion_para_measure_d(rb200_ctx* ctx, const double* mtd_sum, const double* mtd_diff, const double* flags, int V, int R,
                                int extra_dots, const double* r_scale, double delta_r, int r_interp_times,
                                const double* v_scale, double delta_v, int v_interp_times,
                                const double* k_values, int k_rows, int k_cols, double beam_pos_num, double beam_angle_step,
                                int fre_ind, double ele_angle_comp, double ele_angle_sys_err, int mtd_0_num,
                                double* out_r, double* out_v, double* out_ele, int capacity, int* n_out);

/* f4 (DMX script variant): one frame of CW/DMX_SignalProcessing_main_xzr.m:332-426,462-465 for the two monopulse beams.
 *   short pulse (columns [0,n_short)):  filter(fir_taps,1,x) along range per PRT                       (:344-345)
 *   long pulse  (the remaining columns, <= fft_num): ifft(fft(x,fft_num) .* conj(fft(mf,fft_num)))     (:202,348-353)
 *       mf = the (already windowed and normalised) reference, e.g. refData.*kaiser(67,4.5)/norm        (:158-202)
 *   MTD: fft(x .* mtd_window, mtd_fft_num) along slow time, zero-padded, NOT fftshifted                (:414-418)
 *   sum = |L| + |R|, diff = |R| - |L|                                                                 (:421-426)
 *   rows [0, n_blank] and [mtd_fft_num - n_blank, mtd_fft_num) of both sums <- 0 (MTD_0_num = n_blank) (:462-465)
 * left/right: P x n_range column-major split-complex double (imag may be NULL).  Outputs: mtd_fft_num x n_short and
 * mtd_fft_num x fft_num column-major real double (any may be NULL; the *_short ones must be NULL when n_short = 0).
 * The sums feed rb200_execute_cfar_d (n0 = n_blank) and rb200_motion_para_measure_d exactly as in the script.      */
int rb200_dmx_process_z(rb200_ctx* ctx, const double* left_re, const double* left_im, const double* right_re, const double* right_im,
                        int P, int n_range, int n_short, const double* fir_taps, int n_fir,
                        const double* mf_re, const double* mf_im, int n_mf, int fft_num,
                        const double* mtd_window, int mtd_fft_num, int n_blank,
                        double* sum_short, double* diff_short, double* sum_long, double* diff_long);

/* f4: sliding-window CPI assembly with the pulse compression done once (MP/main_produce_dataset_win_xzr.m:24-38:
 * echo_win = [frame N; frame N+1], window i = rows round(i*P/n)+1 ... +P, fun_MTD_produce per window).
 * Pulse compression is per PRT and therefore identical for every window a PRT belongs to; this entry point
 * compresses the P_total rows once and runs MTD + 0-v on n_win windows of win_len rows starting at
 * row_start[i] (0-based).  echo: P_total x R column-major split double; out: n_win consecutive win_len x R
 * column-major real matrices.  Equivalent to n_win calls of rb200_mtd_produce_z on the row slices.        */
int rb200_mtd_produce_windows_z(rb200_ctx* ctx, const double* echo_re, const double* echo_im, int P_total, int R,
                                int win_len, const int32_t* row_start, int n_win, double beta, int zero_v_div, double* out);

/* f1: unpack of DBF-type (data_type 2) PRT payloads: 24-bit little-endian I/Q words, rows of
 * n_channels*6 + one_sample_pad bytes, each PRT payload padded to 64 B (FrameDataRead_xzr.m:111-119,130-135,163).
 * bytes: n_prt concatenated payloads (host).  out_ri: float2 [column][prt][sample] with
 * *n_columns = number of complex columns the M-code's slicing produces (n_channels, or one more when the
 * per-sample padding is 8 bytes).  Values are exact (|v| <= 2^23); 0x800000 decodes to +8388608 as in the M-code. */
int rb200_unpack_dbf24(rb200_ctx* ctx, const uint8_t* bytes, int n_prt, int n_samples, int n_channels,
                       float* out_ri, int* n_columns);

/* f1 (batched): the whole chain on DBF-type frames.  payload = n_cpi x n_prt PRT payloads exactly as they sit in the
 * capture file (FrameDataRead_xzr.m:111-119: n_range rows of 6*n_ch + padding bytes, each PRT padded to 64 B); the
 * complex columns the reference keeps (:130-135,163) become the cfg.n_lanes lanes of the chain (n_lanes must equal
 * that column count, e.g. n_ch = 13 -> 13 lanes).  Everything else as rb200_chain_i16.                          */
int rb200_chain_dbf24(rb200_ctx* ctx, const uint8_t* payload, int n_ch, int n_cpi, float* rdm_out, rb200_det* dets, int* n_det,
                      void* stream);


#ifdef __cplusplus
}
#endif
#endif /* RADAR_B200_H */
