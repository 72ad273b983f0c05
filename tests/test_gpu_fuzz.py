"""Seeded random sweep of the batched chain against the oracle: geometry, lane count, waveform plan, CFAR windows / method /
crop / range stage, iSTC, MTI, zero-velocity divisor and chunking are all drawn per seed, so that every kernel variant
(TMA and non-TMA pulse compression, fused 64- and 256-point MTD + CFAR with compile-time and run-time windows, generic
Stockham, sparse and dense range stage) is crossed with the others.  Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import mcode, synth, vec

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _draw(seed):
    rng = np.random.default_rng(9000 + seed)
    P = int(rng.choice([64, 64, 64, 256, 256, 32, 48, 96, 128]))
    C = int(rng.choice([1, 2, 3, 8, 13, 16, 16, 20]))
    R = int(rng.integers(400, 1500))
    if C >= 13 and P == 256:
        R = int(rng.integers(400, 700))
    B = int(rng.integers(1, 4))
    default_win = rng.random() < 0.5
    ref_v, g_v = (5, 7) if default_win else (int(rng.integers(2, 7)), int(rng.integers(0, 6)))
    ref_r, g_r = (5, 7) if rng.random() < 0.5 else (int(rng.integers(2, 7)), int(rng.integers(0, 6)))
    n0 = int(rng.choice([0, 0, 1, 3]))
    while P - 2 * n0 - 1 < 2 * (ref_v + g_v) + 1:           # keep the velocity axis legal (Function_CFAR1D_sub.m:30-39)
        ref_v, g_v, n0 = max(2, ref_v - 1), max(0, g_v - 1), 0
    cfar = (ref_r, g_r, float(rng.choice([4.0, 5.0, 7.0])), int(rng.integers(0, 2)), ref_v, g_v, float(rng.choice([4.0, 5.0, 7.0])),
            int(rng.integers(0, 2)), n0, int(rng.random() < 0.8))
    mti = int(rng.choice([0, 0, 0, 5, 30])) if P > 40 else 0
    zdiv = int(rng.choice([150, 150, 20]))
    use_stc = rng.random() < 0.3
    chunk = int(rng.integers(1, B + 1))
    ref_name = "refDDCDataMF1" if rng.random() < 0.6 else "refDBFDataMF1"
    lss = bool(rng.random() < 0.3) and R >= 700             # 5-arg segment rule 82 / 242 / rest with the literal pulses
    return dict(P=P, C=C, R=R, B=B, cfar=cfar, mti=mti, zdiv=zdiv, stc=use_stc, chunk=chunk, ref=ref_name, lss=lss)


@pytest.mark.parametrize("seed", range(28))
def test_chain_random_configuration(lib, seed):
    k = _draw(seed)
    P, R, C, B = k["P"], k["R"], k["C"], k["B"]
    ref = mcode.load_ref(k["ref"])
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, ref=ref, n_targets=3, r_lo=20, r_hi=R - 80, seed0=100 * seed)
    stc = synth.s5_stc_curve()[: min(1025, R)] if k["stc"] else None
    if k["lss"]:
        p2, p3 = mcode.load_pulse_literals()
        plan, segs = ("lss_mp", p2, p3), lib.waveforms.segments_mp(R, p2, p3)
    else:
        plan, segs = ("single", ref), lib.waveforms.segments_single(R, ref)
    out = vec.chain(raw, B, P, R, C, plan, k["cfar"], zero_div=k["zdiv"], stc=stc, mti_lag=k["mti"], near_tol=RTOL)
    with lib.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, mti_lag=k["mti"], zero_v_div=k["zdiv"], chunk_cpi=k["chunk"],
                     max_det=1 << 21) as ctx:
        ctx.set_waveform(segs)
        ctx.set_cfar(*k["cfar"])
        if stc is not None:
            ctx.set_stc(stc)
        rdm, dets, n = ctx.chain(raw, B, allow_overflow=False)
    scale = np.abs(out["rdm"]).max()
    assert np.abs(rdm - out["rdm"]).max() <= RTOL * scale, k
    flag, flagv = lib.dets_to_flags(dets, B, C, P, R)
    bad_v = (flagv != out["flagV"]) & ~out["nearV"]
    bad_2 = (flag != out["flag"]) & ~out["near"]
    assert bad_v.sum() == 0 and bad_2.sum() == 0, (k, int(bad_v.sum()), int(bad_2.sum()))
    # the excused (near-threshold) differences are themselves bounded: at most 1e-3 of the cells
    n_diff = int((flagv != out["flagV"]).sum()), int((flag != out["flag"]).sum())
    assert max(n_diff) <= 1e-3 * flag.size, (k, n_diff, flag.size)
