#!/usr/bin/env python3
"""Time one frame of the DMX script variant at the script's own sizes (two beams of 1536 PRT x 566 samples, 62 short + 504 long,
FFT_num 512, mtd_FFT_num 2048) through rb200_dmx_process_z, then executeCFAR on both sums (host doubles in and out).
Content is random (timing only; parity: tests/test_gpu_parity.py::test_dmx_frame_matches_oracle)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radar_signal_process_b200 as rsp
from radar_signal_process_b200 import waveforms

rng = np.random.default_rng(2)
P, n_range, n_short, fft_num, mtd_fft, n0 = 1536, 566, 62, 512, 2048, 27
left = np.rint(rng.normal(0, 100, (P, n_range))) + 1j * np.rint(rng.normal(0, 100, (P, n_range)))
right = np.rint(rng.normal(0, 100, (P, n_range))) + 1j * np.rint(rng.normal(0, 100, (P, n_range)))
ref = waveforms.REF_DDC
mf = ref / np.linalg.norm(ref) * np.kaiser(ref.size, 4.5)
fir = np.array([-9, -7, -2, 10, 27, 40, 42, 24, -13, -57, -89, -86, -30, 77, 220, 364, 471, 511, 471, 364, 220, 77, -30, -86, -89, -57,
                -13, 24, 42, 40, 27, 10, -2, -7, -9], dtype=np.float64)          # filter_coef literal of the script (:146)
win = np.hamming(P)
ctx = rsp.default_context()
for it in range(4):
    t0 = time.perf_counter()
    ss, ds, sl, dl = ctx.dmx_process(left, right, n_short, fir, mf, fft_num, win, mtd_fft, n0)
    t1 = time.perf_counter()
    fs, _ = rsp.executeCFAR(ss, 5, 7, 7.0, 0, 5, 7, 7.0, 0, n0, 1)
    fl, _ = rsp.executeCFAR(sl, 5, 7, 7.0, 0, 5, 7, 7.0, 0, n0, 1)
    t2 = time.perf_counter()
    print("iter %d: dmx_process %.2f ms, 2 x executeCFAR %.2f ms, flags %d + %d" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), int(fs.sum()), int(fl.sum())))
