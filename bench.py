#!/usr/bin/env python3
"""bench.py -- CPI frames/s through unpack -> PC -> MTD -> 0-v -> CFAR on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this framework (CUDA path)
    python bench.py --impl reference --steps K --warmup W      # the reference's CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                         # N > 1: one rank per GPU

A "step" is one pass of the hot path over one batch of ``--cpis`` synthetic CPIs (S3: 64 PRT x 4096
range x 16 lanes, int16 DDC wire format, plan "single" with the captured 67-tap chirp).  Each rank
owns its own batch (weak scaling, no inter-GPU traffic on the hot path); the sparse detection lists
are all-gathered with NCCL after the timed region and reported separately.

Prints ONE JSON line (rank 0).  value = whole-job CPIs/s with inputs resident in HBM, device-timed
with CUDA events on the launching stream, max over ranks.  e2e = same metric through the C-ABI call
with pinned HOST buffers (H2D of the raw samples and D2H of RDM + detections inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Workloads.  S3 is the BASELINE.json headline (configs[2]/[3]); S5 is configs[4]; S1 (configs[0]
# stand-in) goes through the MATLAB-layout entry points and has its own arm below.
WORKLOADS = {
    "S3": dict(P=64, R=4096, C=16, ref="REF_DDC", cfar=(5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1), mti=0, stc=False, cpis=64, distinct=8,
               synth=dict(seed0=1234, r_lo=100, r_hi=3900, exclude=(-1, 0)),
               name="S3 synthetic LFM DDC 4096 range x 64 PRT x 16 lanes, plan single (refDDCDataMF1), CFAR 5/7/T5/GO",
               metric="CPI frames/s (PC->MTD->0v->CFAR, 64 PRT x 4096 range x 16 lanes int16 DDC)",
               kernel="pcw_kernel (K1: int16 unpack + overlap-save pulse compression, warp-private lines; the largest share of the chain's device time)",
               ncu_regex="pcw_kernel",
               kernel_mtd="mtd64_tma_kernel (K2: window + 64-point slow-time FFT + |.| + 0-v + velocity CFAR, register-resident columns)",
               ncu_regex_mtd="mtd64_tma"),
    "S5": dict(P=256, R=16384, C=16, ref="REF_DBF", cfar=(5, 7, 7.0, 0, 5, 7, 7.0, 0, 0, 1), mti=30, stc=True, cpis=4, distinct=1,
               synth=dict(seed0=5000, r_lo=100, r_hi=16384 - 200, exclude=(-3, -2, -1, 0, 1, 2, 3)),
               name="S5 DBF mode 16384 range x 256 PRT x 16 lanes, refDBFDataMF1, iSTC + MTI(30), CFAR 5/7/T7/GO",
               metric="CPI frames/s (iSTC->PC->MTI->MTD->0v->CFAR, 256 PRT x 16384 range x 16 lanes int16)",
               kernel="pcw_kernel (K1: int16 unpack + iSTC + overlap-save pulse compression, warp-private lines)",
               ncu_regex="pcw_kernel",
               kernel_mtd="mtd256_tma_kernel<MTI,2> (K2: persistent, TMA-staged tiles; MTI + window + 256-point slow-time FFT + |.| + 0-v + fused velocity CFAR; the largest share of S5)",
               ncu_regex_mtd="mtd256_tma"),
}
W = dict(WORKLOADS["S3"])                 # the active workload (set in main)
P, R, C = W["P"], W["R"], W["C"]
CELLS = P * R * C
ALG_BYTES_PER_CPI = CELLS * 8          # 4 B int16 I/Q read + 4 B fp32 RDM magnitude written per cell
CFAR = W["cfar"]
METRIC = W["metric"]


def select_workload(name):
    global W, P, R, C, CELLS, ALG_BYTES_PER_CPI, CFAR, METRIC
    W = dict(WORKLOADS[name])
    P, R, C = W["P"], W["R"], W["C"]
    CELLS = P * R * C
    ALG_BYTES_PER_CPI = CELLS * 8
    CFAR = W["cfar"]
    METRIC = W["metric"]


def s5_stc_curve():
    return 30.0 * (1.0 - np.arange(1025) / 1024.0)      # linear 30 -> 0 dB over the first 1025 cells (MP/fun_iSTC.m:6-9)


def synth(n, first_cpi=0, distinct=None):
    from radar_signal_process_b200 import waveforms, workload
    return workload.synth_batch(n, first_cpi=first_cpi, distinct=distinct, P=P, R=R, C=C, ref=getattr(waveforms, W["ref"]), **W["synth"])


def oracle_chain(raw, n, lanes=None, near=False):
    """The checker / CPU baseline: oracle/vec.py on ``n`` CPIs (optionally a subset of lanes); ``near`` adds the
    near-threshold masks the parity check needs (not part of the timed reference path)."""
    from oracle import vec
    from radar_signal_process_b200 import waveforms
    if lanes is not None:
        raw = np.ascontiguousarray(raw[:, :, :, lanes, :])
    nl = raw.shape[3]
    return vec.chain(raw, n, P, R, nl, ("single", getattr(waveforms, W["ref"])), CFAR, stc=s5_stc_curve() if W["stc"] else None,
                     mti_lag=W["mti"], near_tol=1e-4 if near else None)


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 8:
                    self.samples.append(f)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for f in self.samples:
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class NvmlSampler(threading.Thread):
    """SM clock and clock-event reasons read through NVML about every millisecond, only while `active` is set (the
    device-timed region); falls back to the nvidia-smi sampler when NVML is unavailable."""

    def __init__(self, torch_device):
        super().__init__(daemon=True)
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        import torch
        handle = None
        try:
            uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(torch_device.index or 0)
        self.h = handle
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        self.active = threading.Event()
        self.done = threading.Event()
        self.sm, self.bits = [], 0

    def run(self):
        nv = self.nv
        while not self.done.is_set():
            if self.active.is_set():
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.001)

    def stop(self):
        self.done.set()
        self.join(timeout=2)

    def summary(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        reasons = sorted(n for n, bit in names if self.bits & bit)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": 0, "source": "nvml"}
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.sm), "source": "nvml, ~1 ms cadence, device-timed region only"}


# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_cpis, first_cpi=0, workers=None):
    """Time the oracle port of the reference chain (vectorised NumPy/SciPy double precision, all host
    threads for the FFTs) on a bounded sample of the benchmark workload.  Returns (cpis_per_s, seconds, sample text)."""
    from oracle import vec                                   # the checker, executed here only as the CPU baseline
    if workers:
        vec.set_workers(workers)
    if P * R * C > 2 ** 24:                                  # S5: one CPI is 67 M cells -- time 2 of its 16 lanes
        raw = synth(1, first_cpi=first_cpi)
        lanes = [0, C - 1]
        t0 = time.perf_counter()
        oracle_chain(raw, 1, lanes=lanes)
        dt = time.perf_counter() - t0
        return len(lanes) / C / dt, dt, "%d of %d lanes of one CPI through oracle/vec.py (float64, scipy.fft workers=all), %.1f s, scaled" % (len(lanes), C, dt)
    raw = synth(n_cpis, first_cpi=first_cpi)
    t0 = time.perf_counter()
    for i in range(n_cpis):
        oracle_chain(raw[i:i + 1], 1)
    dt = time.perf_counter() - t0
    return n_cpis / dt, dt, "%d CPIs of the workload through oracle/vec.py (float64, scipy.fft workers=all), %.1f s" % (n_cpis, dt)


def cpu_loop_faithful_rate():
    """The loop-faithful transcription (oracle/mcode.py: per-PRT pulse-compression loop with fft(h) recomputed per
    PRT, per-range-cell Doppler loop, per-hit range-CFAR loop; one thread, like a MATLAB script) on ONE lane of one
    S3 CPI, extrapolated x16 lanes.  Returns (cpis_per_s, seconds measured)."""
    from oracle import mcode
    from radar_signal_process_b200 import waveforms, workload
    raw = synth(1)
    x = raw[0, :, :, 0, 0].astype(np.float64) + 1j * raw[0, :, :, 0, 1].astype(np.float64)      # lane 0: P x R
    t0 = time.perf_counter()
    pc = np.zeros(x.shape, dtype=np.complex128)
    L = waveforms.REF_DDC.size
    for i in range(P):                                                   # MTD/fun_lss_pulse_compression.m:36,42,63-65
        pc[i, :] = mcode.fun_pulse_compression(waveforms.REF_DDC, x[i, :])[L - 1:L - 1 + R]
    rdm = mcode.fun_0v_pressing(mcode.fun_Process_MTD(pc, R, P), 150)
    mcode.executeCFAR(rdm, *CFAR)
    dt = time.perf_counter() - t0
    return 1.0 / (dt * C), dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  MATLAB/Octave are absent
    from this image, so this is the oracle port (oracle/vec.py), on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload == "S1":
        return run_s1_reference(args)
    if args.workload == "S2":
        return run_s2_reference(args)
    big = P * R * C > 2 ** 24
    sample = 1 if big else max(1, args.ref_cpis)
    lanes = [0] if big else None                             # S5: one lane of one CPI per step, scaled by the lane count
    raw = synth(sample)

    def step():
        for i in range(sample):
            oracle_chain(raw[i:i + 1], 1, lanes=lanes)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt * ((len(lanes) / C) if lanes else 1.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "CPI/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": W["name"], "cpis_per_step": sample,
                   "impl_note": "oracle port of the M-code (vectorised NumPy/SciPy, float64); MATLAB/Octave absent" +
                                ("; one of %d lanes per step, value scaled by 1/%d" % (C, C) if lanes else "")},
        "cpu_baseline": {"value": value, "unit": "CPI/s", "cores": cores, "kind": "port",
                         "sample": "%d CPI(s)%s per step x %d steps of the workload" % (sample, " (1 lane)" if lanes else "", args.steps)},
        "e2e": {"value": value, "unit": "CPI/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- configuration 1 stand-in (S1): one 1536 x 1031 frame per beam through the MATLAB-layout entry points ----------------
S1_NAME = "S1 frame 1536 PRT x 1031 range, fun_MTD_produce (segments 82/242/707) + crop 691:845 + 0-v/20 + fun_CFARflag (3 x executeCFAR)"
S1_METRIC = "frames/s (fun_MTD_produce + main_cfar.m crop/0-v/fun_CFARflag, 1536 PRT x 1031 range, host doubles in and out)"


def s1_frame(seed):
    rng = np.random.default_rng(seed)
    # column-major like the MATLAB matrix the reference hands to fun_MTD_produce (FrameDataRead_xzr.m:150-156 builds it that way)
    return np.asfortranarray(np.rint(rng.normal(0, 200, (1536, 1031))) + 1j * np.rint(rng.normal(0, 200, (1536, 1031))))


def s1_chain(mod_produce, mod_zero_v, mod_cfar, echo):
    mtd = mod_produce(echo)                                           # MP/main_produce_dataset_win_xzr.m:37-38
    crop = mod_zero_v(np.abs(mtd[690:845, :]), 20)                    # CW/main_cfar.m rows 691:845, CW/fun_0v_pressing.m
    flags = np.zeros_like(crop)
    for c0, c1 in ((0, 82), (82, 318), (318, 868)):                   # fun_CFARflag, CW/main_cfar.m:142-161
        f, _ = mod_cfar(crop[:, c0:c1], 5, 7, 5.0, 0, 5, 7, 5.0, 0, 10, 1)
        flags[:, c0:c1] = f
    return mtd, flags


class _CropView:
    """What fun_MTD_produce_rows returns, indexable like the full matrix for the rows it holds (691:845)."""
    def __init__(self, crop, full):
        self.crop = crop

    def __getitem__(self, key):
        rows, cols = key
        assert rows == slice(690, 845, None)
        return self.crop[:, cols]


def run_s1_reference(args):
    from oracle import mcode, vec
    p2, p3 = mcode.load_pulse_literals()
    echo = s1_frame(1)

    def produce(e):
        return vec.zero_v(vec.process_mtd(vec.lss_pc_mp(e, p2, p3), axis=0), 150, axis=0)

    def cfar(x, *a):
        return vec.execute_cfar(x, *a)[:2]

    def step():
        s1_chain(produce, lambda m, d: vec.zero_v(m, d, axis=0), cfar, echo)

    for _ in range(min(args.warmup, 1)):
        step()
    n = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = time.perf_counter() - t0
    value = n / dt
    print(json.dumps({"impl": "reference", "metric": S1_METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": n,
                      "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": S1_NAME},
                      "cpu_baseline": {"value": value, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                       "sample": "%d frames through oracle/vec.py" % n},
                      "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_s1(args):
    """S1 on the GPU: the reference-facing host API (same call sequence as main_produce_dataset_win_xzr.m + main_cfar.m).
    Inputs and outputs are MATLAB doubles on the host, so every figure here is end to end; `value` repeats it."""
    import torch
    import radar_signal_process_b200 as rsp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    echo = s1_frame(1)
    for _ in range(max(args.warmup, 3)):
        mtd, flags = s1_chain(rsp.fun_MTD_produce, rsp.fun_0v_pressing, rsp.executeCFAR, echo)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mtd, flags = s1_chain(rsp.fun_MTD_produce, rsp.fun_0v_pressing, rsp.executeCFAR, echo)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    peak, peak_src = measured_peak()
    alg = 1536 * 1031 * (16 + 8)                  # complex double in, double out per range-Doppler cell of fun_MTD_produce
    line = {"metric": S1_METRIC, "value": value, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": S1_NAME, "note": "host-double API: value == e2e (no device-resident form of this entry point)"},
            "roofline": {"bound": "hbm", "kernel": "whole fun_MTD_produce call (PCIe + layout conversion + pc_fft + mtd_generic)", "achieved": value * alg / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": value * alg / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 1536 * 1031 * 16 + 155 * 868 * 8 * 4,
                    "d2h_bytes_per_step": 1536 * 1031 * 8 + 155 * 868 * 8 * 7},
            "gpu_launches": None, "detections_per_step": int(flags.sum())}
    # crop-aware variant: fun_MTD_produce_rows(echo, 691, 845) (slow-time transform first, PC on the kept rows only) replaces
    # fun_MTD_produce + the caller's crop; everything downstream is unchanged
    def produce_rows(e):
        full = np.zeros((1536, 1), dtype=np.float64)         # placeholder so that s1_chain's crop indexing stays the caller's
        return _CropView(rsp.fun_MTD_produce_rows(e, 691, 845), full)
    for _ in range(3):
        s1_chain(produce_rows, rsp.fun_0v_pressing, rsp.executeCFAR, echo)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mtd_c, flags_c = s1_chain(produce_rows, rsp.fun_0v_pressing, rsp.executeCFAR, echo)
    torch.cuda.synchronize()
    dtc = time.perf_counter() - t0
    line["e2e_rows"] = {"value": args.steps / dtc, "unit": "frames/s", "ms_per_step": 1e3 * dtc / args.steps,
                        "h2d_bytes_per_step": 1536 * 1031 * 16 + 155 * 868 * 8 * 4, "d2h_bytes_per_step": 155 * 1031 * 8 + 155 * 868 * 8 * 7,
                        "flags_equal_full_path": bool(np.array_equal(flags_c, flags)),
                        "note": "fun_MTD_produce_rows(echo, 691, 845): Doppler first, pulse compression and D2H on the 155 kept rows only"}
    if not args.no_cpu_baseline:
        from oracle import mcode, vec
        p2, p3 = mcode.load_pulse_literals()
        t0 = time.perf_counter()
        want = vec.zero_v(vec.process_mtd(vec.lss_pc_mp(echo, p2, p3), axis=0), 150, axis=0)
        s1_chain(lambda e: want, lambda m, d: vec.zero_v(m, d, axis=0), lambda x, *a: vec.execute_cfar(x, *a)[:2], echo)
        dtc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dtc, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "one frame through oracle/vec.py (float64), %.1f s" % dtc}
        line["parity"] = {"rdm_rel_err": float(np.abs(mtd - want).max() / np.abs(want).max())}
    print(json.dumps(line))


# ---- configuration 2 (S2): MP/main.m simulated target + clutter, 8 PRT x 1024 range, fun_MTD_produce + executeCFAR -------------
S2_NAME = "S2 frame 8 PRT x 1024 range (main.m simulated target + fun_add_clutter), fun_MTD_produce (segments 82/242/700) + executeCFAR 5/7/T5 x 2/1/T5"
S2_METRIC = "frames/s (fun_MTD_produce + executeCFAR, 8 PRT x 1024 range, host doubles in and out)"
S2_CFAR_ARGS = (5, 7, 5.0, 0, 2, 1, 5.0, 0, 0, 1)


def run_s2_reference(args):
    from oracle import mcode, synth
    echo = synth.s2_frame()

    def step():
        mtd = mcode.fun_MTD_produce_mp(echo)
        return mcode.executeCFAR(mtd, *S2_CFAR_ARGS)

    step()
    n = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = time.perf_counter() - t0
    value = n / dt
    print(json.dumps({"impl": "reference", "metric": S2_METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": n,
                      "warmup": 1, "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f64", "data": "synthetic", "config": {"workload": S2_NAME},
                      "cpu_baseline": {"value": value, "unit": "frames/s", "cores": 1, "kind": "port",
                                       "sample": "%d frames through oracle/mcode.py (the loop-faithful transcription)" % n},
                      "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def run_s2(args):
    """S2 on the GPU through the reference-facing host API (MP/main.m:204-215 call sequence): host doubles in and out, so the
    figure is end to end; the frame is tiny (8 x 1024), i.e. this measures call latency, not throughput."""
    import torch
    import radar_signal_process_b200 as rsp
    from oracle import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    echo = np.asfortranarray(synth.s2_frame())

    def step():
        mtd = rsp.fun_MTD_produce(echo)
        return mtd, rsp.executeCFAR(mtd, *S2_CFAR_ARGS)

    for _ in range(max(args.warmup, 3)):
        mtd, (flag, flagv) = step()
    torch.cuda.synchronize()
    n = max(args.steps, 50)
    t0 = time.perf_counter()
    for _ in range(n):
        mtd, (flag, flagv) = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    value = n / dt
    peak, peak_src = measured_peak()
    alg = 8 * 1024 * (16 + 8)
    line = {"metric": S2_METRIC, "value": value, "unit": "frames/s", "n_gpus": 1, "steps": n, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": S2_NAME, "note": "host-double API: value == e2e; latency-bound (about ten kernel launches and four small copies per frame)"},
            "roofline": {"bound": "hbm", "kernel": "whole fun_MTD_produce + executeCFAR call sequence (launch- and copy-latency bound at this size)",
                         "achieved": value * alg / 1e9, "peak": peak, "unit": "GB/s", "frac": value * alg / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 8 * 1024 * 16 + 8 * 1024 * 8, "d2h_bytes_per_step": 8 * 1024 * 8 * 3},
            "gpu_launches": None, "detections_per_step": int(flag.sum())}
    if not args.no_cpu_baseline:
        from oracle import mcode
        t0 = time.perf_counter()
        want = mcode.fun_MTD_produce_mp(echo)
        fo, fvo = mcode.executeCFAR(mtd, *S2_CFAR_ARGS)
        dtc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / dtc, "unit": "frames/s", "cores": 1, "kind": "port",
                                "sample": "one frame through oracle/mcode.py (loop-faithful transcription), %.2f s" % dtc}
        line["parity"] = {"rdm_rel_err": float(np.abs(mtd - want).max() / np.abs(want).max()),
                          "flags_identical_on_gpu_rdm": bool(np.array_equal(flag, fo) and np.array_equal(flagv, fvo))}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
ALL_CPUS = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set()


def bind_to_gpu_numa_node(torch_device):
    """Pin this process to the CPUs NVML reports as local to its GPU, so that the pinned staging buffers (first touch)
    and the copy-submitting thread sit on the GPU's own NUMA node; matters for the end-to-end figure at N > 1."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID((uuid if uuid.startswith("GPU-") else "GPU-" + uuid).encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return "nvml cpu affinity (%d cpus)" % len(os.sched_getaffinity(0))
    except Exception as e:                                    # not fatal: the device-resident figure does not depend on it
        return "unbound (%s)" % type(e).__name__


def ncu_dram_bytes(regex, workload, cpis):
    """--ncu: DRAM bytes per launch of the dominant kernel, measured NOW by a short profiled sub-run of this very script
    (dram__bytes_read.sum + dram__bytes_write.sum, cold cache, averaged over the captured launches)."""
    import csv
    import shutil
    if not shutil.which("ncu"):
        return None, "ncu not found"
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:" + regex, "-s", "1", "-c", "2",
           "--csv", sys.executable, os.path.abspath(__file__), "--workload", workload, "--cpis", str(cpis), "--chunk", str(cpis), "--steps", "1", "--warmup", "1",
           "--no-cpu-baseline", "--e2e-steps", "0", "--no-parity"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout
    except Exception as e:
        return None, "ncu failed: %s" % type(e).__name__
    rows = [r for r in csv.reader(out.splitlines()) if len(r) > 10]
    while rows and "Metric Name" not in rows[0]:      # the profiled script's own JSON line precedes the table on stdout
        rows.pop(0)
    if not rows:
        return None, "ncu produced no rows"
    hdr = rows[0]
    try:
        i_name, i_val, i_unit, i_id = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    except ValueError:
        return None, "unexpected ncu csv"
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_launch = {}
    for r in rows[1:]:
        if r[i_name].startswith("dram__bytes"):
            per_launch.setdefault(r[i_id], 0.0)
            per_launch[r[i_id]] += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
    if not per_launch:
        return None, "no dram metrics in the ncu output"
    return sum(per_launch.values()) / len(per_launch), "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:%s, %d launches, this run" % (regex, len(per_launch))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import radar_signal_process_b200 as rsp
    from radar_signal_process_b200 import distributed as rdist
    from radar_signal_process_b200 import waveforms

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cpis = args.cpis if args.cpis > 0 else W["cpis"]
    distinct = args.distinct if args.distinct > 0 else W["distinct"]
    if args.scaling == "strong":
        # one fixed batch of --cpis CPIs block-partitioned over the ranks
        lo, hi = rdist.shard_range(cpis, rank, world)
        B, first = hi - lo, lo
        total_per_step = cpis
    else:
        # rank r owns global CPIs [r*B, (r+1)*B)  (weak scaling: per-GPU work fixed)
        B, first = cpis, rank * cpis
        total_per_step = world * cpis
    assert B >= 1, "strong scaling: fewer CPIs than ranks"
    raw_np = synth(B, first_cpi=first, distinct=min(distinct, B))
    raw_pin = torch.from_numpy(raw_np).pin_memory()
    raw_dev = raw_pin.to(dev, non_blocking=False)
    rdm_dev = torch.empty((B, C, P, R), dtype=torch.float32, device=dev)
    rdm_pin = torch.empty((B, C, P, R), dtype=torch.float32).pin_memory()

    ctx = rsp.Context(local_rank, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, max_det=args.max_det, chunk_cpi=args.chunk, mti_lag=W["mti"])
    ctx.set_waveform(waveforms.segments_single(R, getattr(waveforms, W["ref"])))
    ctx.set_cfar(*CFAR)
    if W["stc"]:
        ctx.set_stc(s5_stc_curve())
    # a non-default torch stream: its handle goes to the C ABI, so the library's kernels and torch's
    # CUDA events are on the same stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def step_device():
        ctx.chain_enqueue(raw_dev.data_ptr(), B, rdm_dev.data_ptr(), sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timed region ------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler, nvml = None, None
    if rank == 0:
        try:
            nvml = NvmlSampler(dev)
            nvml.start()
        except Exception:
            nvml = None
            sampler = ClockSampler(local_rank)
            sampler.start()
            time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if nvml:
        nvml.active.set()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    # the same step repeated for at least --sustain seconds: a sustained figure next to the K-step one (same clocks record)
    sustained = None
    if args.sustain > 0 and ms > 0:
        reps = max(1, int(np.ceil(args.sustain * 1e3 / ms)))
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        u0.record(stream)
        for _ in range(reps * args.steps):
            step_device()
        u1.record(stream)
        barrier()
        sus_ms = u0.elapsed_time(u1)
        sustained = (reps * args.steps, sus_ms)
    if nvml:
        nvml.active.clear()
        nvml.stop()
    launches_per_step = ctx.last_launch_count()
    dets, n_det = ctx.chain_fetch(allow_overflow=True)

    # ---- parity of the step just timed: CPI 0 of this rank's batch against the oracle (the checker) ---------------
    parity = None
    if rank == 0 and not args.no_parity:
        lanes = [0] if P * R * C > 2 ** 24 else None                    # S5: one lane of the CPI (67 M cells per CPI otherwise)
        out = oracle_chain(raw_np[0:1], 1, lanes=lanes, near=True)
        got = rdm_dev[0].cpu().numpy()
        if lanes is not None:
            got = got[lanes]
        want = out["rdm"][0]
        flag, flagv = rsp.dets_to_flags(dets[dets["cpi"] == 0], 1, C, P, R)
        if lanes is not None:
            flag, flagv = flag[:, lanes], flagv[:, lanes]
        d2 = flag[0] != out["flag"][0]
        dv = flagv[0] != out["flagV"][0]
        parity = {"cpi": 0, "lanes": "all" if lanes is None else lanes, "rdm_rel_err": float(np.abs(got - want).max() / np.abs(want).max()),
                  "rdm_tol": 1e-4, "flags_2d": int(out["flag"][0].sum()), "flags_differ": int(d2.sum()),
                  "flags_differ_unexcused": int((d2 & ~out["near"][0]).sum()), "flags_v_differ": int(dv.sum()),
                  "flags_v_differ_unexcused": int((dv & ~out["nearV"][0]).sum()), "near_count": int(out["near"][0].sum()),
                  "cells": int(want.size), "checker": "oracle/vec.py (float64), near-threshold tolerance 1e-4"}

    # per-kernel durations for the roofline object: the same K steps again with CUDA events around every
    # stage of every chunk (chunks serialised on one stream; the headline `value` above is measured without
    # these events and with chunk pipelining on)
    ctx.set_stage_timing(True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record(stream)
    for _ in range(args.steps):
        step_device()
    s1.record(stream)
    barrier()
    ms_serial = s0.elapsed_time(s1)
    stage_ms, n_chunks, n_stage_cpis = ctx.get_stage_ms()
    ctx.set_stage_timing(False)

    # ---- end-to-end through the C ABI with pinned host buffers ------------------------------------
    dets_pin = torch.empty(args.max_det * 16, dtype=torch.uint8).pin_memory()
    e2e_steps = max(0, min(args.steps, args.e2e_steps))
    e2e_s, e2e_do_s, n_e2e = None, None, 0

    def step_e2e(with_rdm=True):
        st, n = ctx.chain_ptr(raw_pin.data_ptr(), B, rdm_pin.data_ptr() if with_rdm else None, dets_pin.data_ptr(), sptr)
        if st not in (0, 7):
            rsp._binding.raise_for(st, ctx._h)
        return n

    if e2e_steps:
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            n_e2e = step_e2e()
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        # detections only (rdm_out = NULL): the host sees the raw samples go up and only the sparse list come back
        step_e2e(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e(False)
        torch.cuda.synchronize(dev)
        e2e_do_s = time.perf_counter() - t0
    if sampler:
        sampler.stop()

    # ---- sparse detection gather over NCCL, on the device, steady state (median of 10 after one warm-up) ---------
    gather_ms, gather_first_ms, n_gathered = None, None, len(dets)
    if world > 1:
        step_device()
        barrier()
        times = []
        for it in range(11):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            g0.record(stream)
            allrec, counts = rdist.gather_detections_device(ctx, cpi_offset=first, device=dev)
            g1.record(stream)
            torch.cuda.synchronize(dev)
            times.append(g0.elapsed_time(g1))
        gather_first_ms, gather_ms = times[0], statistics.median(times[1:])
        n_gathered = int(counts.sum())
        assert allrec.shape[0] == n_gathered
        # max over ranks of the device times and of the e2e wall times
        vals = [ms, e2e_s or 0.0, e2e_do_s or 0.0, gather_ms, sustained[1] if sustained else 0.0]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, gather_ms = float(t[0]), float(t[3])
        if e2e_s is not None:
            e2e_s, e2e_do_s = float(t[1]), float(t[2])
        if sustained:
            sustained = (sustained[0], float(t[4]))

    if rank == 0:
        peak, peak_src = measured_peak()
        value = total_per_step * args.steps / (ms * 1e-3)
        # the roofline object describes the DOMINANT kernel of this workload: the stage with the largest measured share
        # (S3: K1 pulse compression; S5: the P = 256 Doppler + velocity-CFAR kernel)
        dom = max(("pc", "mtd"), key=lambda k: stage_ms[k])
        dom_kernel = W["kernel"] if dom == "pc" else W["kernel_mtd"]
        dom_regex = W["ncu_regex"] if dom == "pc" else W["ncu_regex_mtd"]
        pc_ms_per_launch = stage_ms[dom] / max(n_chunks, 1)
        cpis_per_launch = n_stage_cpis / max(n_chunks, 1)
        achieved = ALG_BYTES_PER_CPI * cpis_per_launch / (pc_ms_per_launch * 1e-3) / 1e9 if pc_ms_per_launch > 0 else None
        stage_total = sum(stage_ms.values())
        traffic, traffic_src = None, "not measured in this run (pass --ncu)"
        if args.ncu and world == 1:
            # same CPIs per launch as the line above, so that `traffic` and `algorithmic_bytes_per_launch` describe the same launch
            per_launch, traffic_src = ncu_dram_bytes(dom_regex, args.workload, max(int(round(cpis_per_launch)), 1))
            traffic = per_launch
        line = {
            "metric": METRIC, "value": value, "unit": "CPI/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": W["name"],
                       "cpis_per_step_per_gpu": B, "cpis_per_step_total": total_per_step, "distinct_cpis": min(distinct, B), "chunk_cpi": args.chunk,
                       "l2": "inputs larger than L2 (%.0f MiB raw + %.0f MiB RDM per step per GPU)" % (B * CELLS * 4 / 2 ** 20, B * CELLS * 4 / 2 ** 20),
                       "parallelism": "cpi-shard x%d (%s), no hot-path collective" % (world, args.scaling), "host_binding": numa},
            "hbm_gbs_chain": value / world * ALG_BYTES_PER_CPI / 1e9,
            "hbm_frac_chain": value / world * ALG_BYTES_PER_CPI / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": dom_kernel,
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_CPI * cpis_per_launch,
                         "ms_per_launch": pc_ms_per_launch,
                         "serialised_ms_per_step": ms_serial / args.steps,
                         "stage_share": {k: (v / stage_total if stage_total else None) for k, v in stage_ms.items()},
                         "stage_us_per_cpi": {k: 1e3 * v / max(n_stage_cpis, 1) for k, v in stage_ms.items()}},
            "gpu_launches": launches_per_step * args.steps,
            "detections_per_step": n_det, "gather_ms": gather_ms, "gather_first_ms": gather_first_ms, "detections_gathered": n_gathered,
            "gather": "device-side: counts all_gather (1 x int32 per rank) + all_gather of exactly max(count) 16-byte records, median of 10" if world > 1 else None,
            "clocks": nvml.summary() if nvml else (sampler.summary() if sampler else None),
        }
        if e2e_s:
            line["e2e"] = {"value": total_per_step * e2e_steps / e2e_s, "unit": "CPI/s", "h2d_bytes_per_step": B * CELLS * 4,
                           "d2h_bytes_per_step": B * CELLS * 4 + 16 * min(n_e2e, args.max_det), "steps": e2e_steps}
            line["e2e_dets_only"] = {"value": total_per_step * e2e_steps / e2e_do_s, "unit": "CPI/s", "h2d_bytes_per_step": B * CELLS * 4,
                                     "d2h_bytes_per_step": 16 * min(n_e2e, args.max_det), "steps": e2e_steps,
                                     "note": "rdm_out = NULL: only the sparse detection list returns to the host"}
        if sustained:
            line["sustained"] = {"value": total_per_step * sustained[0] / (sustained[1] * 1e-3), "unit": "CPI/s", "steps": sustained[0],
                                 "seconds": sustained[1] * 1e-3}
        if parity is not None:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            try:
                os.sched_setaffinity(0, ALL_CPUS)            # the CPU baseline may use every host core again
            except Exception:
                pass
            cps, dt, sample_txt = cpu_reference_rate(args.cpu_cpis)
            line["cpu_baseline"] = {"value": cps, "unit": "CPI/s", "cores": os.cpu_count() or 1, "kind": "port", "sample": sample_txt}
            if args.workload == "S3":
                loop_cps, loop_dt = cpu_loop_faithful_rate()
                line["cpu_baseline"]["loop_faithful"] = {"value": loop_cps, "unit": "CPI/s", "cores": 1,
                                                         "sample": "oracle/mcode.py (M-code loop structure) on 1 of 16 lanes of one CPI, %.1f s, extrapolated x16" % loop_dt}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="S3", choices=["S3", "S5", "S1", "S2"],
                    help="S3: BASELINE headline (default); S5: DBF long-CPI sweep with iSTC + MTI; S1 / S2: config-1 / config-2 frame through the MATLAB-layout API")
    ap.add_argument("--cpis", type=int, default=0, help="CPIs per step per GPU (0 = workload default: S3 64 -> 1 GiB of raw input, S5 4)")
    ap.add_argument("--distinct", type=int, default=0, help="distinct synthetic CPIs generated (tiled up to --cpis; 0 = workload default)")
    ap.add_argument("--chunk", type=int, default=0, help="CPIs per PC->MTD->CFAR pass (0 = library default)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --cpis is the total batch, block-partitioned over ranks")
    ap.add_argument("--max-det", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustain", type=float, default=1.0, help="seconds of the additional sustained loop (0 = off)")
    ap.add_argument("--cpu-cpis", type=int, default=8, help="CPIs timed for the cpu_baseline object")
    ap.add_argument("--ref-cpis", type=int, default=2, help="CPIs per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of CPI 0 of the timed step")
    ap.add_argument("--ncu", action="store_true", help="fill roofline.traffic from a short ncu sub-run of the dominant kernel")
    args = ap.parse_args()
    if args.workload not in ("S1", "S2"):
        select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "S1":
        run_s1(args)
    elif args.workload == "S2":
        run_s2(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
