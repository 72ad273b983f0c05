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

P, R, C = 64, 4096, 16
CELLS = P * R * C
ALG_BYTES_PER_CPI = CELLS * 8          # 4 B int16 I/Q read + 4 B fp32 RDM magnitude written per cell
CFAR = (5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)
K1_DRAM_BYTES_PER_CPI = (134.590208e6 + 215.659008e6) / 8.0     # ncu, cold cache, see profiles/r01j_ncu_full_final.txt
METRIC = "CPI frames/s (PC->MTD->0v->CFAR, 64 PRT x 4096 range x 16 lanes int16 DDC)"


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
    threads for the FFTs) on ``n_cpis`` CPIs of the benchmark workload.  Returns (cpis_per_s, seconds)."""
    from oracle import vec                                   # the checker, executed here only as the CPU baseline
    from radar_signal_process_b200 import waveforms, workload
    if workers:
        vec.set_workers(workers)
    raw = workload.synth_batch(n_cpis, first_cpi=first_cpi)
    t0 = time.perf_counter()
    for i in range(n_cpis):
        vec.chain(raw[i:i + 1], 1, P, R, C, ("single", waveforms.REF_DDC), CFAR)
    dt = time.perf_counter() - t0
    return n_cpis / dt, dt


def cpu_loop_faithful_rate():
    """The loop-faithful transcription (oracle/mcode.py: per-PRT pulse-compression loop with fft(h) recomputed per
    PRT, per-range-cell Doppler loop, per-hit range-CFAR loop; one thread, like a MATLAB script) on ONE lane of one
    S3 CPI, extrapolated x16 lanes.  Returns (cpis_per_s, seconds measured)."""
    from oracle import mcode
    from radar_signal_process_b200 import waveforms, workload
    raw = workload.synth_batch(1)
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
    sample = max(1, args.ref_cpis)
    from oracle import vec
    from radar_signal_process_b200 import waveforms, workload
    raw = workload.synth_batch(sample)

    def step():
        for i in range(sample):
            vec.chain(raw[i:i + 1], 1, P, R, C, ("single", waveforms.REF_DDC), CFAR)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "CPI/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "S3 synthetic LFM DDC 4096 range x 64 PRT x 16 lanes, plan single (refDDCDataMF1), CFAR 5/7/T5/GO",
                   "cpis_per_step": sample, "impl_note": "oracle port of the M-code (vectorised NumPy/SciPy, float64); MATLAB/Octave absent"},
        "cpu_baseline": {"value": value, "unit": "CPI/s", "cores": cores, "kind": "port",
                         "sample": "%d CPIs per step x %d steps of the S3 workload" % (sample, args.steps)},
        "e2e": {"value": value, "unit": "CPI/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import radar_signal_process_b200 as rsp
    from radar_signal_process_b200 import distributed as rdist
    from radar_signal_process_b200 import waveforms, workload

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

    B = args.cpis
    # rank r owns global CPIs [r*B, (r+1)*B)  (weak scaling: per-GPU work fixed)
    raw_np = workload.synth_batch(B, first_cpi=rank * B, distinct=min(args.distinct, B))
    raw_pin = torch.from_numpy(raw_np).pin_memory()
    raw_dev = raw_pin.to(dev, non_blocking=False)
    rdm_dev = torch.empty((B, C, P, R), dtype=torch.float32, device=dev)
    rdm_pin = torch.empty((B, C, P, R), dtype=torch.float32).pin_memory()

    ctx = rsp.Context(local_rank, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, max_det=args.max_det, chunk_cpi=args.chunk)
    ctx.set_waveform(waveforms.segments_single(R, waveforms.REF_DDC))
    ctx.set_cfar(*CFAR)
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
    if nvml:
        nvml.active.clear()
        nvml.stop()
    ms = e0.elapsed_time(e1)
    launches_per_step = ctx.last_launch_count()
    dets, n_det = ctx.chain_fetch(allow_overflow=True)
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
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def step_e2e():
        st, n = ctx.chain_ptr(raw_pin.data_ptr(), B, rdm_pin.data_ptr(), dets_pin.data_ptr(), sptr)
        if st not in (0, 7):
            rsp._binding.raise_for(st, ctx._h)
        return n

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n_e2e = step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if sampler:
        sampler.stop()

    # ---- sparse detection gather over NCCL (outside the hot path, timed separately) ----------------
    gather_ms = None
    n_gathered = len(dets)
    if world > 1:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        allrec, counts = rdist.gather_detections(dets, args.max_det, cpi_offset=rank * B, device=dev)
        g1.record(stream)
        torch.cuda.synchronize(dev)
        gather_ms = g0.elapsed_time(g1)
        n_gathered = int(counts.sum())
        # max over ranks of the device time and of the e2e wall time
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = measured_peak()
        total_cpis = world * B * args.steps
        value = total_cpis / (ms * 1e-3)
        pc_ms_per_launch = stage_ms["pc"] / max(n_chunks, 1)
        cpis_per_launch = n_stage_cpis / max(n_chunks, 1)
        achieved = ALG_BYTES_PER_CPI * cpis_per_launch / (pc_ms_per_launch * 1e-3) / 1e9 if pc_ms_per_launch > 0 else None
        stage_total = sum(stage_ms.values())
        line = {
            "metric": METRIC, "value": value, "unit": "CPI/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "S3 synthetic LFM DDC 4096 range x 64 PRT x 16 lanes, plan single (refDDCDataMF1), CFAR 5/7/T5/GO",
                       "cpis_per_step_per_gpu": B, "distinct_cpis": min(args.distinct, B), "chunk_cpi": args.chunk,
                       "l2": "inputs larger than L2 (%.0f MiB raw + %.0f MiB RDM per step)" % (B * CELLS * 4 / 2 ** 20, B * CELLS * 4 / 2 ** 20),
                       "parallelism": "cpi-shard x%d, no hot-path collective" % world, "host_binding": numa},
            "hbm_gbs_chain": value / world * ALG_BYTES_PER_CPI / 1e9,
            "hbm_frac_chain": value / world * ALG_BYTES_PER_CPI / 1e9 / peak,
            "roofline": {"bound": "hbm", "kernel": "pc_fft_tma_kernel (K1: int16 unpack + overlap-save pulse compression; 60 % of the chain's device time)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, from the ncu --set full
                         # capture profiles/r01j_ncu_full_final.txt (134.6 MB read + 215.7 MB written per 8-CPI launch)
                         "traffic": K1_DRAM_BYTES_PER_CPI * cpis_per_launch, "traffic_source": "profiles/r01j_ncu_full_final.txt",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_CPI * cpis_per_launch,
                         "ms_per_launch": pc_ms_per_launch,
                         "serialised_ms_per_step": ms_serial / args.steps,
                         "stage_share": {k: (v / stage_total if stage_total else None) for k, v in stage_ms.items()},
                         "stage_us_per_cpi": {k: 1e3 * v / max(n_stage_cpis, 1) for k, v in stage_ms.items()}},
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "CPI/s", "h2d_bytes_per_step": B * CELLS * 4,
                    "d2h_bytes_per_step": B * CELLS * 4 + 16 * min(n_e2e, args.max_det), "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "detections_per_step": n_det, "gather_ms": gather_ms, "detections_gathered": n_gathered,
            "clocks": nvml.summary() if nvml else (sampler.summary() if sampler else None),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                os.sched_setaffinity(0, ALL_CPUS)            # the CPU baseline may use every host core again
            except Exception:
                pass
            cps, dt = cpu_reference_rate(args.cpu_cpis)
            loop_cps, loop_dt = cpu_loop_faithful_rate()
            line["cpu_baseline"] = {"value": cps, "unit": "CPI/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "%d CPIs of the S3 workload through oracle/vec.py (float64, scipy.fft workers=all), %.1f s" % (args.cpu_cpis, dt),
                                    "loop_faithful": {"value": loop_cps, "unit": "CPI/s", "cores": 1,
                                                      "sample": "oracle/mcode.py (M-code loop structure) on 1 of 16 lanes of one CPI, %.1f s, extrapolated x16" % loop_dt}}
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
    ap.add_argument("--cpis", type=int, default=64, help="CPIs per step per GPU (64 -> 1 GiB of raw input)")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic CPIs generated (tiled up to --cpis)")
    ap.add_argument("--chunk", type=int, default=0, help="CPIs per PC->MTD->CFAR pass (0 = library default)")
    ap.add_argument("--max-det", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-cpis", type=int, default=8, help="CPIs timed for the cpu_baseline object")
    ap.add_argument("--ref-cpis", type=int, default=2, help="CPIs per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
