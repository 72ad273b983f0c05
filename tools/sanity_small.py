#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel family, meant to be run under compute-sanitizer
(memcheck / racecheck, one tool per gpurun call):  compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radar_signal_process_b200 as rsp
from radar_signal_process_b200 import waveforms as W
from radar_signal_process_b200 import workload

rng = np.random.default_rng(0)


def rc(*s):
    return rng.normal(size=s) + 1j * rng.normal(size=s)


# fused path (TMA PC + TMA MTD64 + sparse range stage), 16 lanes, two chunks over two slots
P, R, C, B = 64, 768, 16, 3
raw = workload.synth_batch(B, P=P, R=R, C=C, n_targets=2, r_lo=20, r_hi=R - 80)
with rsp.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, chunk_cpi=2) as ctx:
    ctx.set_waveform(W.segments_single(R, W.REF_DDC))
    ctx.set_cfar(5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)
    rdm, dets, n = ctx.chain(raw, B)
    print("fused chain:", n, "detections, rdm max %.3e" % rdm.max())

# generic path: odd lane count, lss plan (FIR + two MF segments), MTI, non-default CFAR, generic Doppler length
P, R, C, B = 48, 1031, 5, 2
raw = rng.integers(-300, 300, size=(B, P, R, C, 2), dtype=np.int16)
with rsp.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, mti_lag=7, chunk_cpi=1) as ctx:
    ctx.set_waveform(W.segments_mp(R, W.PULSE2, W.PULSE3))
    ctx.set_cfar(4, 3, 4.0, 1, 4, 3, 4.0, 1, 1, 1)
    ctx.set_stc(np.linspace(20, 0, 100))
    rdm, dets, n = ctx.chain(raw, B, allow_overflow=True)
    print("generic chain:", n, "detections")

# P = 256 shared-memory Doppler path
P, R, C, B = 256, 512, 2, 1
raw = rng.integers(-300, 300, size=(B, P, R, C, 2), dtype=np.int16)
with rsp.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, mti_lag=30) as ctx:
    ctx.set_waveform(W.segments_single(R, W.REF_DBF))
    ctx.set_cfar(5, 7, 7.0, 0, 5, 7, 7.0, 0, 0, 1)
    rdm, dets, n = ctx.chain(raw, B, allow_overflow=True)
    print("P=256 chain:", n, "detections")

# MATLAB-layout entry points
echo = np.rint(100 * rc(40, 1031))
m = rsp.fun_MTD_produce(echo)
pc = rsp.fun_lss_pulse_compression(echo, 0, W.pulse1_mp(), W.PULSE2, W.PULSE3)
y = rsp.fun_pulse_compression(rc(700), rc(900))          # 4096-sample tile
y2 = rsp.fun_pulse_compression(rc(5000), rc(64))         # time-domain fallback
mt = rsp.fun_Process_MTD(rc(155, 33), 33, 155)
z = rsp.fun_0v_pressing(np.abs(mt), 20)
f, fv = rsp.executeCFAR(z, 5, 7, 3.0, 0, 5, 7, 3.0, 0, 2, 1)
g = rsp.Function_CFAR1D_sub_fixCells(z, 5, 7, 2.0, 0, [1, 3], [1, 17, 33])
print("matlab api ok", m.shape, pc.shape, y.shape, y2.shape, mt.shape, f.sum(), g.sum())
rsp.shutdown()
