"""Pin the oracle against every known-answer artefact the reference ships."""
import hashlib
import json
import os

import numpy as np
import scipy.signal.windows as ssw

from oracle import mcode

G = os.path.join(os.path.dirname(__file__), "golden")


def test_fixture_digests():
    want = json.load(open(os.path.join(G, "sha256.json")))
    for fn, dig in want.items():
        if fn.endswith(".npz"):
            z = np.load(os.path.join(G, fn))
            h = hashlib.sha256()
            for key in sorted(z.files):
                h.update(key.encode())
                h.update(np.ascontiguousarray(z[key]).tobytes())
            assert h.hexdigest() == dig, fn
        else:
            a = np.load(os.path.join(G, fn))
            assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == dig, fn


def test_kaiser_matches_matlab_fixture():
    # MP/kaiser_win.mat == MATLAB kaiser(1536, 8), the window of MP/fun_Process_MTD.m:13-15
    k = np.load(os.path.join(G, "kaiser_win_1536.npy"))
    assert k.shape == (1536,)
    assert np.max(np.abs(mcode.kaiser(1536, 8) - k)) < 5e-15
    assert np.max(np.abs(ssw.kaiser(1536, 8, sym=True) - k)) < 5e-15


def test_kaiser_small():
    assert np.array_equal(mcode.kaiser(1, 8), np.ones(1))
    for n in (2, 8, 64, 155, 256):
        assert np.allclose(mcode.kaiser(n, 8), ssw.kaiser(n, 8, sym=True), rtol=0, atol=2e-15)


def test_literals_shapes_and_values():
    p2, p3 = mcode.load_pulse_literals()
    assert p2.shape == (75,) and p3.shape == (160,)
    # first / last literals of MP/fun_MTD_produce.m:54-59
    assert p2[0] == 8.73485e-18 - 0.010182j and p2[-1] == -0.00962958 - 0.00330821j
    assert p3[0] == 0.00347321 - 0.00601577j and p3[-1] == 0.00680575 + 0.00139083j
    assert mcode.FILTER_COEF_INT.shape == (35,) and mcode.FILTER_COEF_INT.max() == 511
    assert np.array_equal(mcode.FILTER_COEF_INT, mcode.FILTER_COEF_INT[::-1])


def test_ref_chirps():
    d = mcode.load_ref("refDDCDataMF1")
    b = mcode.load_ref("refDBFDataMF1")
    assert d.shape == (67,) and b.shape == (67,)
    assert np.array_equal(d, np.round(d))          # integer-valued captured chirp
    assert abs(np.abs(d).max() - 7588.2289106220305) < 1e-9
    assert d[0] == 84 - 11j and d[1] == -1246 + 3277j


def test_pulse1_is_seven_points():
    p1 = mcode.pulse1_mp()                          # MP/fun_MTD_produce.m:51 comment: 7 points
    assert p1.shape == (7,)


def test_mround_half_away():
    assert mcode.mround(155 / 2) == 78
    assert mcode.mround(2.5) == 3 and mcode.mround(-2.5) == -3 and mcode.mround(0.4) == 0


def test_zero_v_rows_table():
    # SURVEY 7.4-1 table
    assert mcode.zero_v_rows(1536, 150) == (758, 778)
    assert mcode.zero_v_rows(64, 150) == (32, 32)
    assert mcode.zero_v_rows(8, 150) == (4, 4)
    assert mcode.zero_v_rows(256, 150) == (126, 130)
    assert mcode.zero_v_rows(155, 20) == (70, 86)


def test_grpdelay_is_17():
    b = mcode.FILTER_COEF_INT / 511.0
    assert mcode.grpdelay_mean_round(b) == 17       # MTD/fun_lss_pulse_compression.m:47


def test_matlab_colon_follows_the_published_algorithm():
    """``a:d:b`` (MTD/fun_MTD_produce.m:62-63 builds the ideal LFM time axes with it): element count with the end-point
    tolerance, right end snapped to b, symmetric fill from both ends, exact mid-point for an even interval count --
    MathWorks' published COLONOP replication, not ``a + d*(0:n)``."""
    from oracle import mcode
    from radar_signal_process_b200 import waveforms
    assert np.array_equal(mcode.matlab_colon(0, 1, 5), np.arange(6.0))
    assert np.array_equal(mcode.matlab_colon(1, 2, 10), np.array([1.0, 3, 5, 7, 9]))
    assert mcode.matlab_colon(5, 1, 1).size == 0 and mcode.matlab_colon(0, 0, 1).size == 0
    v = mcode.matlab_colon(0, 0.1, 1)
    assert v.size == 11 and v[-1] == 1.0 and v[5] == 0.5                 # end snapped to b, mid-point (a + c) / 2
    assert v[10 - 3] == 1.0 - 3 * 0.1 and v[3] == 3 * 0.1                # filled from both ends: 0.7 from the right (naive: 0.7000000000000001)
    ts = 1 / 25e6
    for tao in (0.28e-6, 8e-6, 28e-6):
        t = mcode.matlab_colon(-tao / 2, ts, tao / 2 - ts)
        assert t.size == int(round(tao / ts)) and t[0] == -tao / 2 and t[-1] == tao / 2 - ts
        n = t.size - 1
        kk = np.arange((n + 1) // 2)                                          # right half, the exact mid-point of an even n excluded
        assert np.array_equal(t[n - kk], (tao / 2 - ts) - kk * ts) and np.array_equal(t[kk], -tao / 2 + kk * ts)
        if n % 2 == 0:
            assert t[n // 2] == (-tao / 2 + (tao / 2 - ts)) / 2
        assert np.array_equal(t, waveforms._colon(-tao / 2, ts, tao / 2 - ts))     # the host mirror makes the same choice
