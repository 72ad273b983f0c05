"""Synthetic benchmark workload: LFM point-target echoes in the int16 DDC wire format.

Input generation only (configuration S3): per CPI and lane, ``n_targets``
point targets (range ~U[r_lo, r_hi], Doppler bin ~U over the CPI excluding the listed bins), each a
replica of the 67-tap captured chirp scaled to a post-integration SNR ~U[15,30] dB, plus rounded
Gaussian noise on I and Q, saturated to int16.  Layout ``[cpi][prt][range][lane][I,Q]`` as produced
per PRT by FrameDataRead_xzr.m:150-156.  Both bench.py arms consume this generator.
"""
import numpy as np

from . import waveforms


def _kaiser(n, beta):
    if n == 1:
        return np.ones(1)
    a = (n - 1) / 2.0
    r = (np.arange(n) - a) / a
    return np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / np.i0(beta)


def synth_cpi(cpi, P=64, R=4096, C=16, ref=None, n_targets=8, noise_sigma=64.0, seed0=1234, r_lo=100, r_hi=3900,
              exclude=(-1, 0)):
    ref = waveforms.REF_DDC if ref is None else np.asarray(ref)
    rng = np.random.default_rng(seed0 + cpi)
    L = ref.shape[0]
    w = _kaiser(P, 8.0)
    a = rng.normal(0.0, noise_sigma, size=(P, R, C, 2))
    np.rint(a, out=a)
    p = np.arange(P)
    e_ref = float(np.sum(np.abs(ref) ** 2))
    bins = [k for k in range(-(P // 2), P - P // 2) if k not in exclude]
    targets = []
    for lane in range(C):
        for _ in range(n_targets):
            r0 = int(rng.integers(r_lo, r_hi + 1))
            k = int(bins[int(rng.integers(0, len(bins)))])
            snr_db = float(rng.uniform(15.0, 30.0))
            amp = np.sqrt(10.0 ** (snr_db / 10.0) * 2.0 * noise_sigma ** 2 * float((w ** 2).sum()) / (e_ref * float(w.sum()) ** 2))
            ph = np.exp(2j * np.pi * k * p / P)
            nv = min(L, R - r0)
            echo = amp * ph[:, None] * ref[None, :nv]
            a[:, r0:r0 + nv, lane, 0] += echo.real
            a[:, r0:r0 + nv, lane, 1] += echo.imag
            targets.append((lane, r0, k, snr_db))
    np.rint(a, out=a)
    return a.clip(-32768, 32767).astype(np.int16), targets


def synth_batch(n_cpi, first_cpi=0, distinct=None, **kw):
    """``n_cpi`` CPIs; if ``distinct`` < n_cpi the first ``distinct`` CPIs are generated and tiled."""
    distinct = n_cpi if distinct is None else min(distinct, n_cpi)
    base = np.stack([synth_cpi(first_cpi + c, **kw)[0] for c in range(distinct)], axis=0)
    if distinct == n_cpi:
        return base
    reps = (n_cpi + distinct - 1) // distinct
    return np.concatenate([base] * reps, axis=0)[:n_cpi]
