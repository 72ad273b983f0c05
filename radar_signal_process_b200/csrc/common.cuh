// common.cuh -- shared device-side parameter blocks and small helpers for libradar_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb {

constexpr int kMaxSegs = 8;

__host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember the largest size configured on each.
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kernel, size_t bytes, size_t (&done)[64]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    if (bytes <= done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done[dev] = bytes;
    return e;
}

// One waveform segment in correlation form:
//   y[n] = sum_k xin[n + k - pre] * conj(t[k]),  n in [0, out_len),  xin = 0 outside [0, in_len)
//   out[out_start + ((n - rot) mod out_len)] = y[n]
// MF leading edge: pre = 0, rot = 0 (MP/fun_lss_pulse_compression.m:36-37);
// FIR delayed:     pre = L-1, t = conj(flip(b)) (MP/...:25-31);
// FIR grpdelay:    same with rot = round(mean(grpdelay(b))) (MTD/fun_lss_pulse_compression.m:47-51).
struct PcSegDev {
    int in_start, in_len, out_start, out_len;
    int rot, V, pre, h_off;   // V = outputs per overlap-save tile, h_off = offset of the spectrum in hperm
    int n_taps, t_off;        // time-domain taps (direct fallback): offset into taps table
};

struct PcParams {
    const void* in;         // int16 wire [group][range][lane][2]  or  float2 planar [line][range]
    float2* out;            // float2 planar [line][R]
    const float2* hperm;    // digit-reversed reference spectra, (scale/NT) * conj(FFT(taps))
    const float2* tw;       // exp(-2*pi*i*m/NT), m = 0..NT-1
    const float* gain;      // iSTC linear gain per range cell (nullable)
    const int2* tiles;      // (segment, tile index) per blockIdx.y
    PcSegDev segs[kMaxSegs];
    int R;                  // samples per input PRT line
    int R_out;              // elements per output line
    int C;                  // lanes interleaved in the wire format
    int P;                  // PRTs per CPI (wire output line mapping)
    int n_lines;            // planar: number of lines
    // RB200_SPLIT (producer / consumer kernels on disjoint SMs): pcw_kernel publishes its progress, one increment per finished
    // warp task in cpi_done[cpi of this launch], and sets *started when its first CTA runs (both null otherwise)
    int* cpi_done;
    int* started;
};

// Range segments of the CFAR (fun_CFARflag, CW/main_cfar.m:142-161): the range stage never looks across a segment border
// and columns outside every segment produce no detections.  n = 0: one segment covering the whole axis.
struct CfarSegs {
    int n;
    int lo[4], hi[4];       // 0-based [lo, hi)
};
__host__ __device__ __forceinline__ bool cfar_seg_of(const CfarSegs& s, int r, int R, int* lo, int* hi) {
    if (s.n == 0) { *lo = 0; *hi = R; return r >= 0 && r < R; }
    for (int i = 0; i < s.n; ++i)
        if (r >= s.lo[i] && r < s.hi[i]) { *lo = s.lo[i]; *hi = s.hi[i]; return true; }
    return false;
}

struct CfarParams {
    int V, R;               // full RDM size
    int v_lo, v_hi;         // 0-based tested rows [v_lo, v_hi)  (n0+1 .. V-n0)
    int ref_r, guard_r, meth_r;
    int ref_v, guard_v, meth_v;
    int range_stage;
    int max_det;
    int n_lanes;            // slabs per CPI (detection record: cpi = slab / n_lanes, lane = slab % n_lanes)
    int cpi0;               // CPI index of slab 0 (chunked batches)
    CfarSegs segs;
};

struct MtdParams {
    const float2* in;       // planar [slab][prt][range]
    float* out;             // [slab][v][range]
    const float* window;    // P
    const float2* tw;       // exp(-2*pi*i*m/P)
    int P;
    int in_ld, out_ld;      // elements per input line / per output row
    int cols;               // columns to process (Len_PRT)
    int mti_lag;            // 0 = off
    int zv_lo, zv_hi;       // 0-based inclusive rows to zero (zv_lo > zv_hi = none)
    // generic Stockham only
    int n_stages;
    int radix[16];
    int in_rows;            // rows present in the input (0 = P); rows in_rows..P-1 are zero padding (fft(x, P) with P > size)
    int no_shift;           // 1: natural DFT order, no fftshift (CW/DMX_SignalProcessing_main_xzr.m:414)
    float2* out_c;          // non-null: write the COMPLEX spectrum of output rows crop_lo..crop_hi (0-based, after the shift) as
    int crop_lo, crop_hi;   //           planar lines [row - crop_lo][out_ld] instead of magnitudes (zero-velocity rows as zeros)
    // optional fused velocity-axis CFAR (mtd_fast_kernel only): magnitudes of the CTA's tile are kept in shared
    // memory and every thread decides R consecutive rows of its column; hits go to dets / vmask like cfar_v_kernel
    int cfar_on;
    CfarParams cf;
    float t_v;
    void* dets;
    int* det_count;
    uint32_t* vmask;
    int* err_flag;
    // P = 256: SMs of the device for the persistent TMA-fed kernel (0 = use the one-tile-per-CTA kernel), RB200_NO_TMA_MTD
    int n_sms;
    int no_tma;
};

struct rb200_det_fwd;
// P = 64 register-column kernel (mtd64_kernel.cu): window / keep factors travel in the parameter block
struct Mtd64Params {
    const float2* in;       // planar [slab][prt][range]
    float* out;             // [slab][v][range]
    int in_ld, out_ld, cols;
    float win[64];          // Kaiser window
    float keep[64];         // 0 for zero-velocity rows, 1 elsewhere (per output row)
    // fused velocity-axis CFAR
    int meth_v;
    float tv_over_ref;      // T_V / refCells_V
    void* dets;             // rb200_det list (velocity hits)
    int* det_count;
    unsigned long long* colmask;   // [slab][cols_ld]: bit v set = velocity hit at (v, r)
    int cols_ld;
    int max_det, n_lanes, cpi0;
    CfarSegs segs;          // velocity hits in columns outside every segment are dropped
    // RB200_SPLIT: a tile of CPI c (of this launch) is fetched only when wait_done[c] >= wait_target (null = no waiting)
    const int* wait_done;
    int wait_target;
    int* err_flag;          // set to 2 when that wait times out
};

// fused persistent chain for P = 64, 16 channels (chain64_kernel.cu)
struct Chain64Params {
    PcParams pc;            // in = raw wire samples of the whole batch; out unused (ring below); R, R_out, segs, hperm, tw, tiles, gain
    Mtd64Params m;          // out = RDM of the whole batch, dets / det_count / colmask for the whole batch, cpi0 = 0
    int n_cpi;
    int n_tiles;            // overlap-save tiles per PRT line group
    int n_pc_items;         // per CPI: 64 * n_tiles
    int n_mtd_items;        // per CPI: 16 * ceil(R / 256)
    float2* ring;           // pulse-compressed intermediate, ring_slots CPIs of [lane][prt][range]
    int ring_slots;
    size_t ring_stride;     // float2 elements per CPI slot
    int* work_counter;      // zeroed before launch
    int* pc_done;           // [n_cpi], zeroed before launch
    int* mtd_done;          // [n_cpi], zeroed before launch
    int* err_flag;
};

// single-pass chain for P = 64, 16 interleaved lanes (onepass_kernel.cu); the input arrives through a tensor map of the
// de-interleaved lane planes (a separate kernel argument)
struct OnePassParams {
    float* rdm;                     // [cpi][lane][v][range] of this launch
    const float2* hperm;            // reference spectrum of the single MF segment, position q*16 + k <-> bin q + 16*k, times scale/256
    const float2* tw;               // tw[k*16 + u] = exp(-2*pi*i*u*k/256)
    unsigned long long* colmask;    // [cpi*16 + lane][R]: bit v = velocity hit at (v, r)
    void* dets;                     // rb200_det list of velocity hits (per-slot scratch list)
    int* det_count;
    int* err_flag;
    int max_det;
    int R, V, n_tiles, n_cpi, cpi0;
    int meth_v;
    float tv_over_ref;
    int dbg;                        // timing experiments only (RB200_OP_DBG): 1 skip the Doppler column work, 2 skip the transforms
    float win[64];                  // Kaiser window
    float keep[64];                 // 0 for zero-velocity rows, 1 elsewhere (per output row)
    CfarSegs segs;
};

}  // namespace rb
