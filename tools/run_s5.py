#!/usr/bin/env python3
"""Time configuration S5 (DBF mode, 256 PRT x 16384 range x 16 lanes, iSTC + MTI) at full size on one GPU.
Usage: python tools/run_s5.py [n_cpi]   (prints device ms per CPI; parity is covered by tests/test_gpu_parity.py)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radar_signal_process_b200 as rsp
from radar_signal_process_b200 import waveforms, workload

P, R, C = 256, 16384, 16
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
one, _ = workload.synth_cpi(0, P=P, R=R, C=C, ref=waveforms.REF_DBF, seed0=5000, r_lo=100, r_hi=R - 200, exclude=(-3, -2, -1, 0, 1, 2, 3))
raw = np.stack([one] * n)
stc = 30.0 * (1.0 - np.arange(1025) / 1024.0)
with rsp.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=n, mti_lag=30, max_det=1 << 20) as ctx:
    ctx.set_waveform(waveforms.segments_single(R, waveforms.REF_DBF))
    ctx.set_cfar(5, 7, 7.0, 0, 5, 7, 7.0, 0, 0, 1)
    ctx.set_stc(stc)
    ctx.set_stage_timing(True)
    for it in range(3):
        t0 = time.perf_counter()
        rdm, dets, nd = ctx.chain(raw, n, want_rdm=False, allow_overflow=True)
        wall = time.perf_counter() - t0
        print("iter %d: device %.3f ms/CPI (%d CPIs), wall %.1f ms, %d detections, %d launches" %
              (it, ctx.last_device_ms() / n, n, 1e3 * wall, nd, ctx.last_launch_count()))
        ms, nch, ncp = ctx.get_stage_ms()
        print("   stage ms per CPI:", {k: round(v / max(ncp, 1), 3) for k, v in ms.items()})
