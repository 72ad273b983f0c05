#!/usr/bin/env python3
"""Time BASELINE config 1 (one 1536 x 1031 frame through fun_MTD_produce, then the main_cfar.m crop + zero-velocity
+ three-segment executeCFAR) through the MATLAB-layout entry points: host doubles in, host doubles out, everything a MEX
call would do.  Content is random (timing only; parity is covered by tests/test_gpu_parity.py::test_fun_MTD_produce_S1)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radar_signal_process_b200 as rsp

rng = np.random.default_rng(1)
P, R = 1536, 1031
echo = np.rint(rng.normal(0, 200, (P, R))) + 1j * np.rint(rng.normal(0, 200, (P, R)))
for it in range(4):
    t0 = time.perf_counter()
    mtd = rsp.fun_MTD_produce(echo)                                   # MP/main_produce_dataset_win_xzr.m:37-38
    t1 = time.perf_counter()
    crop = np.abs(mtd[690:845, :])                                    # CW/main_cfar.m: rows 691:845
    crop = rsp.fun_0v_pressing(crop, 20)                              # CW/fun_0v_pressing.m (divisor 20)
    flags = np.zeros_like(crop)
    for c0, c1 in ((0, 82), (82, 318), (318, 868)):                   # fun_CFARflag, CW/main_cfar.m:142-161
        f, _ = rsp.executeCFAR(crop[:, c0:c1], 5, 7, 5.0, 0, 5, 7, 5.0, 0, 10, 1)
        flags[:, c0:c1] = f
    t2 = time.perf_counter()
    print("iter %d: fun_MTD_produce %.2f ms, crop + 0-v + 3 x executeCFAR %.2f ms, %d flags" %
          (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), int(flags.sum())))
