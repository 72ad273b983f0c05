"""Vectorised double-precision twin of ``oracle.mcode`` (TEST ORACLE / timed CPU baseline).

Same maths as the loop-faithful transcription, restructured as whole-array NumPy/SciPy calls so
that the large configurations (S3: 64 x 4096 x 16, S5: 256 x 16384 x 16) finish in seconds.
``tests/test_oracle_equivalence.py`` asserts it equal to ``oracle.mcode`` on small shapes
(bit-identical CFAR flags; PC/RDM to ~1e-12 relative because the FFT length differs).

Closed forms used (each proved against the transcription by the tests):

* matched-filter segment: ``out[n] = sum_k x[n+k] conj(ref[k])``, n = 0..M-1, x = 0 beyond M
  (== ``ifft(fft(x,n).*fft(conj(flip(ref)),n))(L:L+M-1)``, MP/fun_pulse_compression.m:4-22 +
  MP/fun_lss_pulse_compression.m:36-37);
* FIR segment: ``y[n] = sum_k b[k] x[n-k] / 1.2`` (MP/fun_lss_pulse_compression.m:21-26);
* MTD: ``|fftshift(fft(x .* kaiser(P,8)))|`` along slow time (MP/fun_Process_MTD.m:13-29);
* executeCFAR == dense formulation: F[v,c]=1 iff exists r in {c-1,c,c+1}: V[v,r] and Rg[v,c] and
  c = first-argmax of x[v,.] over {r-1,r,r+1} & Rg  (CW/executeCFAR.m:36-89).
"""
import os

import numpy as np
import scipy.fft as sfft

from . import mcode
from .mcode import MatlabError, FILTER_COEF_INT, kaiser, zero_v_rows  # noqa: F401

_WORKERS = int(os.environ.get("ORACLE_FFT_WORKERS", "0")) or os.cpu_count() or 1


def set_workers(n):
    global _WORKERS
    _WORKERS = max(1, int(n))


# ----------------------------------------------------------------------------------------------
# A1 unpack, whole CPI batch at once
# ----------------------------------------------------------------------------------------------
def unpack_wire(raw, n_cpi, P, R, C):
    """int16 wire ``[cpi][prt][range][channel][I,Q]`` -> complex128 ``[cpi][lane][prt][range]``.

    Per-PRT rule of FrameDataRead_xzr.m:150-156 applied to every PRT (payload only, no framing).
    """
    a = np.asarray(raw).view(np.int16).reshape(n_cpi, P, R, C, 2).astype(np.float64)
    z = a[..., 0] + 1j * a[..., 1]
    return np.ascontiguousarray(z.transpose(0, 3, 1, 2))


# ----------------------------------------------------------------------------------------------
# A3/A4 pulse compression
# ----------------------------------------------------------------------------------------------
def mf_rows(x, ref):
    """Row-wise ``out[n] = sum_k x[..., n+k] conj(ref[k])`` (n < M, x zero-extended)."""
    x = np.asarray(x)
    ref = np.asarray(ref).ravel()
    M = x.shape[-1]
    L = ref.shape[0]
    if M == 0:
        return np.zeros(x.shape, dtype=np.complex128)
    n = sfft.next_fast_len(M + L - 1, real=False)
    X = sfft.fft(x, n, axis=-1, workers=_WORKERS)
    H = sfft.fft(np.conj(ref[::-1]), n)
    y = sfft.ifft(X * H, n, axis=-1, workers=_WORKERS)
    return y[..., L - 1:L - 1 + M]


def fir_rows(x, b):
    """Row-wise causal FIR ``y[n] = sum_k b[k] x[n-k]``, truncated to the input length."""
    x = np.asarray(x)
    b = np.asarray(b, dtype=np.float64).ravel()
    M = x.shape[-1]
    y = np.zeros(x.shape, dtype=np.complex128)
    for k in range(min(b.shape[0], M)):
        y[..., k:] += b[k] * x[..., :M - k]
    return y


def lss_pc_mp(echo, pulse2, pulse3):
    """5-arg MP/fun_lss_pulse_compression.m (segments 82/242/rest, FIR left delayed)."""
    echo = np.asarray(echo)
    n = echo.shape[-1]
    if np.size(pulse2) != 75 or np.size(pulse3) != 160:
        raise MatlabError("MATLAB:subsassigndimmismatch", "5-arg API needs 75/160-sample pulses")
    if n < 324:
        raise MatlabError("MATLAB:badsubscript", "Index in position 2 exceeds array bounds")
    out = np.zeros(echo.shape, dtype=np.complex128)
    b = FILTER_COEF_INT / FILTER_COEF_INT.max()
    out[..., :82] = fir_rows(echo[..., :82], b) / 1.2
    out[..., 82:324] = mf_rows(echo[..., 82:324], pulse2)
    out[..., 324:] = mf_rows(echo[..., 324:], pulse3)
    return out


def lss_pc_mtd(echo, pulse2, pulse3, p1, p2, p3):
    """9-arg MTD/fun_lss_pulse_compression.m (segment sizes as arguments, FIR group-delay shifted)."""
    echo = np.asarray(echo)
    n = echo.shape[-1]
    p1, p2, p3 = int(p1), int(p2), int(p3)
    if n < p1 + p2 or p3 > n - p1 - p2:
        raise MatlabError("MATLAB:badsubscript", "Index exceeds array bounds")
    out = np.zeros(echo.shape, dtype=np.complex128)
    b = FILTER_COEF_INT / FILTER_COEF_INT.max()
    d = mcode.grpdelay_mean_round(b)
    out[..., :p1] = np.roll(fir_rows(echo[..., :p1], b) / 1.2, -d, axis=-1)
    out[..., p1:p1 + p2] = mf_rows(echo[..., p1:p1 + p2], pulse2)
    out[..., p1 + p2:p1 + p2 + p3] = mf_rows(echo[..., p1 + p2:], pulse3)[..., :p3]
    return out


def single_pc(echo, ref):
    """One matched-filter segment spanning the whole PRT (9-arg alignment rule, MTD/...:63-65)."""
    return mf_rows(echo, ref)


# ----------------------------------------------------------------------------------------------
# A5/A6/A11/A12
# ----------------------------------------------------------------------------------------------
def process_mtd(x, beta=8.0, window=None, axis=-2):
    """|fftshift(fft(x .* w))| along the slow-time axis (default: second-to-last = PRT)."""
    x = np.asarray(x)
    P = x.shape[axis]
    w = kaiser(P, beta) if window is None else np.asarray(window, dtype=np.float64)
    shape = [1] * x.ndim
    shape[axis] = P
    X = sfft.fft(x * w.reshape(shape), axis=axis, workers=_WORKERS)
    return np.abs(np.roll(X, P // 2, axis=axis))


def zero_v(rdm, div=150, axis=-2):
    rdm = np.array(rdm, copy=True)
    lo, hi = zero_v_rows(rdm.shape[axis], div)
    if lo < 1 or hi > rdm.shape[axis]:
        raise MatlabError("MATLAB:badsubscript", "Index in position 1 is invalid")
    sl = [slice(None)] * rdm.ndim
    sl[axis] = slice(lo - 1, hi)
    rdm[tuple(sl)] = 0
    return rdm


def istc(echo, stc_ini):
    echo = np.asarray(echo)
    n = echo.shape[-1]
    stc = np.zeros(n)
    s = np.asarray(stc_ini, dtype=np.float64).ravel()
    if s.size > n:
        raise MatlabError("MATLAB:sizeDimensionsMustMatch", "STC curve longer than the PRT")
    stc[: s.size] = s
    return echo * (10.0 ** (stc / 20.0))


def mti(x, lag=30, axis=-2):
    x = np.asarray(x)
    x = np.moveaxis(x, axis, 0)
    out = np.zeros_like(x)
    P = x.shape[0]
    if P > lag:
        out[: P - lag] = x[lag:] - x[: P - lag]
    return np.moveaxis(out, 0, axis)


# ----------------------------------------------------------------------------------------------
# A8 dense 1-D CA-CFAR along the last axis, with decision margins
# ----------------------------------------------------------------------------------------------
def cfar1d_last(data, ref, guard, T, method, want_margin=False):
    """Function_CFAR1D_sub applied along the last axis of ``data`` (any leading shape).

    Window sums are accumulated left to right exactly like ``mean(..., 2)`` in the transcription,
    so flags are bit-identical to ``mcode.Function_CFAR1D_sub``.
    Returns flags (float64 0/1) and, optionally, the relative margin |x - T*mu| / max(|T*mu|, tiny).
    """
    data = np.asarray(data, dtype=np.float64)
    N = data.shape[-1]
    ref = int(ref)
    guard = int(guard)
    if N < 2 * (ref + guard):
        raise MatlabError("MATLAB:badsubscript", "Index exceeds array bounds (axis shorter than 2*(ref+guard))")
    lead = data.shape[:-1]
    y = np.arange(N)
    sumL = np.zeros(lead + (N,))
    sumR = np.zeros(lead + (N,))
    okL = (y - guard - ref) >= 0
    okR = (y + guard + ref) <= N - 1
    for j in range(ref):
        # left window cells in increasing column order: y-guard-ref+j
        idx = y - guard - ref + j
        sumL[..., okL] += data[..., idx[okL]]
        idx = y + guard + 1 + j
        sumR[..., okR] += data[..., idx[okR]]
    meanL = sumL / ref
    meanR = sumR / ref
    mL = np.where(okL, meanL, meanR)
    mR = np.where(okR, meanR, meanL)
    mu = np.maximum(mL, mR) if method == 0 else np.minimum(mL, mR)
    thr = mu * T
    flags = (data >= thr).astype(np.float64)
    if not want_margin:
        return flags
    with np.errstate(over="ignore"):
        margin = np.abs(data - thr) / np.maximum(np.abs(thr), np.finfo(np.float64).tiny)
    return flags, margin


# ----------------------------------------------------------------------------------------------
# A7 executeCFAR, dense formulation
# ----------------------------------------------------------------------------------------------
def execute_cfar(mtd, refR, saveR, T_R, methR, refV, saveV, T_V, methV, n0, rflag, near_tol=None):
    """Dense executeCFAR on ``mtd[..., V, R]`` (leading dims = batch).  Returns (flag, flagV[, near]).

    ``near`` (if ``near_tol`` is given) marks final-flag cells whose value may legitimately differ in
    a float32 implementation: a contributing velocity/range decision lies within ``near_tol``
    (relative) of its threshold, or two candidate amplitudes tie within ``near_tol``.
    """
    mtd = np.asarray(mtd, dtype=np.float64)
    V, R = mtd.shape[-2:]
    n0 = int(n0)
    lo, hi = n0 + 2, V - n0
    if lo < 1 or hi > V or hi < lo:
        raise MatlabError("MATLAB:badsubscript", "Index in position 1 exceeds array bounds")
    used = mtd[..., lo - 1:hi, :]
    ut = np.swapaxes(used, -1, -2)
    if near_tol is None:
        Vf = np.swapaxes(cfar1d_last(ut, refV, saveV, T_V, methV), -1, -2)
        mV = None
    else:
        Vf, mV = cfar1d_last(ut, refV, saveV, T_V, methV, want_margin=True)
        Vf = np.swapaxes(Vf, -1, -2)
        mV = np.swapaxes(mV, -1, -2)
    flagV = np.zeros(mtd.shape)
    flagV[..., lo - 1:hi, :] = Vf
    flag = np.zeros(mtd.shape)
    near = None
    if rflag:
        if not Vf.any():
            resR = np.zeros(used.shape)
            if near_tol is not None:
                nearu = _dilate(mV <= near_tol, 1)
        else:
            if near_tol is None:
                Rg = cfar1d_last(used, refR, saveR, T_R, methR)
                mR = None
            else:
                Rg, mR = cfar1d_last(used, refR, saveR, T_R, methR, want_margin=True)
            resR = _combine(used, Vf > 0, Rg > 0)
            if near_tol is not None:
                tie = _near_ties(used, near_tol)
                nearu = _dilate(mV <= near_tol, 1) | _dilate(mR <= near_tol, 2) | _dilate(tie, 2)
        flag[..., lo - 1:hi, :] = resR
    else:
        flag = flagV.copy()
        if near_tol is not None:
            nearu = mV <= near_tol
    if near_tol is None:
        return flag, flagV
    near = np.zeros(mtd.shape, dtype=bool)
    near[..., lo - 1:hi, :] = nearu
    nearV = np.zeros(mtd.shape, dtype=bool)
    nearV[..., lo - 1:hi, :] = mV <= near_tol
    return flag, flagV, near, nearV


def _shift(a, s, fill):
    """b[..., c] = a[..., c+s] (fill outside)."""
    out = np.full(a.shape, fill, dtype=a.dtype)
    R = a.shape[-1]
    if s == 0:
        out[...] = a
    elif s > 0:
        if s < R:
            out[..., :R - s] = a[..., s:]
    else:
        if -s < R:
            out[..., -s:] = a[..., :R + s]
    return out


def _combine(x, Vf, Rg):
    """For every velocity hit (v,r): winner = first argmax of x over {r-1,r,r+1} & Rg; set it."""
    neg = -np.inf
    cand = np.stack([_shift(np.where(Rg, x, neg), s, neg) for s in (-1, 0, 1)], axis=-1)
    has = np.isfinite(cand).any(axis=-1) & Vf
    win = np.argmax(cand, axis=-1) - 1          # first maximum, offsets -1,0,+1
    out = np.zeros(x.shape)
    idx = np.nonzero(has)
    col = idx[-1] + win[idx]
    out[idx[:-1] + (col,)] = 1.0
    return out


def _dilate(mask, k):
    out = mask.copy()
    for s in range(1, k + 1):
        out |= _shift(mask, s, False) | _shift(mask, -s, False)
    return out


def _near_ties(x, tol):
    t = np.zeros(x.shape, dtype=bool)
    for s in (1, 2):
        xs = _shift(x, s, np.nan)
        with np.errstate(invalid="ignore"):
            t |= np.abs(x - xs) <= tol * np.maximum(np.abs(x), np.abs(xs))
    return t


def cfar_flag_segments(mtd, cfar_args, segments=((1, 82), (83, 318), (319, 868))):
    """fun_CFARflag of CW/main_cfar.m:142-161 on top of the dense executeCFAR."""
    mtd = np.asarray(mtd, dtype=np.float64)
    out = np.zeros(mtd.shape)
    for a, b in segments:
        f, _ = execute_cfar(mtd[..., a - 1:b], *cfar_args)
        out[..., a - 1:b] = f
    return out


# ----------------------------------------------------------------------------------------------
# whole chain on the int16 wire format (the benchmark workload)
# ----------------------------------------------------------------------------------------------
def dbf_weighting(x, W):
    """``sig_C * W.'`` of FrameDataRead_xzr.m:158 on x[cpi, channel, prt, range] -> [cpi, beam, prt, range]."""
    return np.einsum("bc,icpr->ibpr", np.asarray(W, dtype=np.complex128), x)


def chain(raw, n_cpi, P, R, C, plan, cfar, beta=8.0, zero_div=150, stc=None, mti_lag=0, near_tol=None, dbf=None):
    """unpack -> [iSTC] -> PC -> [MTI] -> MTD -> 0-v -> executeCFAR for every (cpi, lane).

    ``plan`` is ("single", ref) | ("lss_mp", pulse2, pulse3) | ("lss_mtd", pulse2, pulse3, p1, p2, p3).
    ``cfar`` = (refR, saveR, T_R, methR, refV, saveV, T_V, methV, n0, rflag).
    Returns dict(rdm[cpi,lane,V,R], flag, flagV[, near, nearV]).
    """
    x = unpack_wire(raw, n_cpi, P, R, C)
    if dbf is not None:
        x = dbf_weighting(x, dbf)
    return chain_lanes(x, plan, cfar, beta=beta, zero_div=zero_div, stc=stc, mti_lag=mti_lag, near_tol=near_tol)


def chain_lanes(x, plan, cfar, beta=8.0, zero_div=150, stc=None, mti_lag=0, near_tol=None):
    """The chain downstream of the unpack, on complex lanes x[cpi, lane, prt, range] (e.g. decoded DBF-type beams)."""
    if stc is not None:
        x = istc(x, stc)
    if plan[0] == "single":
        pc = single_pc(x, plan[1])
    elif plan[0] == "lss_mp":
        pc = lss_pc_mp(x, plan[1], plan[2])
    elif plan[0] == "lss_mtd":
        pc = lss_pc_mtd(x, *plan[1:])
    else:
        raise ValueError(plan[0])
    if mti_lag:
        pc = mti(pc, mti_lag)
    rdm = zero_v(process_mtd(pc, beta), zero_div)
    res = execute_cfar(rdm, *cfar, near_tol=near_tol)
    out = {"pc": pc, "rdm": rdm, "flag": res[0], "flagV": res[1]}
    if near_tol is not None:
        out["near"], out["nearV"] = res[2], res[3]
    return out
