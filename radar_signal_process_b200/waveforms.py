"""Waveform plans: how the reference's pulse-compression calls map onto rb200_segment lists.

Host-side logic only (argument checking and plan construction, exactly what a MEX gateway does
before it calls the C ABI); all signal arithmetic happens in libradar_b200.so on the GPU.
"""
import math

import numpy as np

from . import _literals
from ._binding import (ALIGN_DELAYED, ALIGN_GRPDELAY, ALIGN_LEADING_EDGE, ERR_DIM_MISMATCH, ERR_INDEX, SEG_FIR, SEG_MF,
                       MatlabDimensionError, MatlabIndexError)

# MP/fun_lss_pulse_compression.m:21-22
FILTER_COEF = np.array(_literals.FILTER_COEF, dtype=np.float64)
FIR_TAPS = FILTER_COEF / FILTER_COEF.max()
# MP/fun_MTD_produce.m:54-60
PULSE2 = np.array(_literals.PULSE2_REAL, dtype=np.float64) + 1j * np.array(_literals.PULSE2_IMAG, dtype=np.float64)
PULSE3 = np.array(_literals.PULSE3_REAL, dtype=np.float64) + 1j * np.array(_literals.PULSE3_IMAG, dtype=np.float64)


def _colon(a, d, b):
    """MATLAB ``a:d:b`` (MathWorks' published COLONOP replication): the element count tolerates rounding of the end
    point, the right end snaps to ``b`` when it is within 2 eps, and the vector is filled symmetrically from both ends
    (``a + k d`` up to the middle, ``c - k d`` down from the right end, the exact mid-point ``(a + c) / 2`` for even n), so
    the last bits differ from the naive ``a + d * (0:n)``."""
    eps = np.finfo(float).eps
    if not (np.isfinite(a) and np.isfinite(d) and np.isfinite(b)):
        return np.array([np.nan])
    if d == 0 or (a < b and d < 0) or (b < a and d > 0):
        return np.zeros(0)
    tol = 2.0 * eps * max(abs(a), abs(b))
    sig = 1.0 if d > 0 else -1.0
    if a == math.floor(a) and d == 1:
        n = int(math.floor(b) - a)
    elif a == math.floor(a) and d == math.floor(d):
        q = math.floor(a / d)
        r = a - q * d
        n = int(math.floor((b - r) / d) - q)
    else:
        n = int(math.floor((b - a) / d + 0.5))
        if sig * (a + n * d - b) > tol:
            n -= 1
    if n < 0:
        return np.zeros(0)
    c = a + n * d
    if sig * (c - b) > -tol:
        c = b
    out = np.zeros(n + 1)
    k = np.arange(n // 2 + 1)
    out[k] = a + k * d
    out[n - k] = c - k * d
    if n % 2 == 0:
        out[n // 2] = (a + c) / 2
    return out


def pulse1_mp():
    """MP/fun_MTD_produce.m:24-27,47,51."""
    ts = 1 / 25e6
    tao1 = 0.28e-6
    return np.sin(2 * np.pi * _colon(-tao1 / 2, ts, tao1 / 2 - ts) + np.pi / 2)


def ideal_pulses(params):
    """MTD/fun_MTD_produce.m:37-38,45-51,61-69 (K2 = -B/tao2, K3 = +B/tao3)."""
    fs = float(params["fs"])
    ts = 1 / fs
    B = float(params["B"])
    tao = [float(t) for t in np.ravel(params["tao"])[:3]]
    t1 = _colon(-tao[0] / 2, ts, tao[0] / 2 - ts)
    t2 = _colon(-tao[1] / 2, ts, tao[1] / 2 - ts)
    t3 = _colon(-tao[2] / 2, ts, tao[2] / 2 - ts)
    return (np.sin(2 * np.pi * t1 + np.pi / 2),
            np.exp(1j * 2 * np.pi * (0.5 * (-B / tao[1]) * t2 ** 2)),
            np.exp(1j * 2 * np.pi * (0.5 * (B / tao[2]) * t3 ** 2)))


def _seg(in_start, in_len, out_start, out_len, kind, align, taps, scale=1.0):
    return dict(in_start=int(in_start), in_len=int(in_len), out_start=int(out_start), out_len=int(out_len),
                kind=kind, align=align, taps=np.asarray(taps, dtype=np.complex128).ravel(), scale=float(scale))


def segments_mp(n, pulse2, pulse3):
    """5-arg MP/fun_lss_pulse_compression.m: hard-coded 82/242/rest, output offsets 75/160 (:6-8,36-37)."""
    pulse2 = np.asarray(pulse2).ravel()
    pulse3 = np.asarray(pulse3).ravel()
    if n < 324:
        raise MatlabIndexError(ERR_INDEX, "fun_lss_pulse_compression: Index in position 2 exceeds array bounds (PRT shorter than 324)")
    if pulse2.size != 75:
        raise MatlabDimensionError(ERR_DIM_MISMATCH, "fun_lss_pulse_compression: pulse2 must have 75 samples (signal_PC_02(75:end) -> 242 columns)")
    if pulse3.size != 160:
        raise MatlabDimensionError(ERR_DIM_MISMATCH, "fun_lss_pulse_compression: pulse3 must have 160 samples (signal_PC_03(160:end))")
    return [
        _seg(0, 82, 0, 82, SEG_FIR, ALIGN_DELAYED, FIR_TAPS, 1 / 1.2),
        _seg(82, 242, 82, 242, SEG_MF, ALIGN_LEADING_EDGE, pulse2),
        _seg(324, n - 324, 324, n - 324, SEG_MF, ALIGN_LEADING_EDGE, pulse3),
    ]


def segments_mtd(n, pulse2, pulse3, p1, p2, p3):
    """9-arg MTD/fun_lss_pulse_compression.m: sizes as arguments, FIR realigned by its group delay (:47-65)."""
    p1, p2, p3 = int(p1), int(p2), int(p3)
    if p1 < 0 or p2 < 0 or p3 < 0 or n < p1 + p2:
        raise MatlabIndexError(ERR_INDEX, "fun_lss_pulse_compression: Index in position 2 exceeds array bounds")
    if p3 > n - p1 - p2:
        raise MatlabIndexError(ERR_INDEX, "fun_lss_pulse_compression: Index exceeds the number of array elements (point_prt3 too large)")
    return [
        _seg(0, p1, 0, p1, SEG_FIR, ALIGN_GRPDELAY, FIR_TAPS, 1 / 1.2),
        _seg(p1, p2, p1, p2, SEG_MF, ALIGN_LEADING_EDGE, pulse2),
        _seg(p1 + p2, n - p1 - p2, p1 + p2, p3, SEG_MF, ALIGN_LEADING_EDGE, pulse3),
    ]


def segments_single(n, ref):
    """One matched-filter segment over the whole PRT (benchmark plan 'single')."""
    return [_seg(0, n, 0, n, SEG_MF, ALIGN_LEADING_EDGE, ref)]

# MP/refDDCDataMF1.mat / MP/refDBFDataMF1.mat (variable refData, 67 taps; load sites
# CW/DMX_SignalProcessing_main_xzr.m:157-159)
REF_DDC = np.array(_literals.REF_DDC_REAL, dtype=np.float64) + 1j * np.array(_literals.REF_DDC_IMAG, dtype=np.float64)
REF_DBF = np.array(_literals.REF_DBF_REAL, dtype=np.float64) + 1j * np.array(_literals.REF_DBF_IMAG, dtype=np.float64)
