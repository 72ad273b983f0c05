"""The MEX gateways executed on the GPU through the shim runtime, against the CPU oracle: this is the
reference-facing boundary a MATLAB/Octave session would hit (same names, argument order, 1-based indices,
column-major doubles)."""
import numpy as np
import pytest

from mexharness import Mex, MexError
from oracle import mcode, synth, vec

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-4):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * np.max(np.abs(b))


def _rc(rng, *s):
    return rng.normal(size=s) + 1j * rng.normal(size=s)


def test_mex_fun_pulse_compression():
    rng = np.random.default_rng(1)
    s0, x = _rc(rng, 67), _rc(rng, 300)
    got = Mex("fun_pulse_compression")(s0, x)
    assert got.shape == (1, 366)
    _close(got[0], mcode.fun_pulse_compression(s0, x))
    got = Mex("fun_pulse_compression")(s0.real, x.real)           # real inputs: imag treated as 0
    _close(got[0], mcode.fun_pulse_compression(s0.real, x.real))


def test_mex_fun_lss_pulse_compression_both_generations():
    rng = np.random.default_rng(2)
    p2, p3 = mcode.load_pulse_literals()
    echo = np.rint(100 * _rc(rng, 5, 1031))
    got = Mex("fun_lss_pulse_compression")(echo, 0, mcode.pulse1_mp(), p2, p3)
    _close(got, mcode.fun_lss_pulse_compression_mp(echo, None, p2, p3))
    with pytest.raises(MexError) as e:
        Mex("fun_lss_pulse_compression")(echo, 0, 1.0, p2[:70], p3)
    assert e.value.ident == "radar_b200:pc:dimensionMismatch"
    params = dict(fs=25e6, B=20e6, tao=[0.16e-6, 8e-6, 28e-6], point_prt=[3404, 228, 723, 2453])
    pulse1, pulse2, pulse3 = mcode.ideal_pulses_mtd(params)
    echo = np.rint(100 * _rc(rng, 3, 3404))
    got = Mex("fun_lss_pulse_compression")(echo, params, 0, pulse1, pulse2, pulse3, 228, 723, 2453)
    _close(got, mcode.fun_lss_pulse_compression_mtd(echo, pulse1, pulse2, pulse3, 228, 723, 2453))
    with pytest.raises(MexError) as e:
        Mex("fun_lss_pulse_compression")(echo, params, 0, pulse1, pulse2, pulse3, 228, 723, 2454)
    assert e.value.ident == "radar_b200:pc:indexOutOfRange"


def test_mex_fun_Process_MTD_and_0v():
    rng = np.random.default_rng(3)
    x = 100 * _rc(rng, 64, 21)
    got = Mex("fun_Process_MTD")(x, 21, 64)
    want = mcode.fun_Process_MTD(x, 21, 64)
    _close(got, want)
    with pytest.raises(MexError) as e:
        Mex("fun_Process_MTD")(x, 22, 64)
    assert e.value.ident == "radar_b200:mtd:indexOutOfRange"
    with pytest.raises(MexError) as e:
        Mex("fun_Process_MTD")(x, 21, 63)
    assert e.value.ident == "radar_b200:mtd:dimensionMismatch"
    assert np.array_equal(Mex("fun_0v_pressing")(want), mcode.fun_0v_pressing(want, 150))
    m155 = rng.rayleigh(1, size=(155, 7))
    assert np.array_equal(Mex("fun_0v_pressing_cw")(m155), mcode.fun_0v_pressing(m155, 20))


def test_mex_fun_MTD_produce_1arg_and_2arg():
    echo = synth.s2_frame()
    got = Mex("fun_MTD_produce")(echo)
    _close(got, mcode.fun_MTD_produce_mp(echo))
    rng = np.random.default_rng(4)
    params = dict(prt=1e-4, prf=1e4, prtNum=16, fs=25e6, deltaR=6.0, fc=9.4e9, wavelength=0.0319, B=20e6,
                  tao=[0.16e-6, 8e-6, 28e-6], point_prt=[3404, 228, 723, 2453], debug=dict(show_PC=0, show_FFT=0, graph=0))
    echo = np.rint(100 * _rc(rng, 16, 3404))
    got = Mex("fun_MTD_produce")(echo, params)
    _close(got, mcode.fun_MTD_produce_mtd(echo, params))
    bad = dict(params, point_prt=[3404, 228, 723, 2500])
    with pytest.raises(MexError) as e:
        Mex("fun_MTD_produce")(echo, bad)
    assert e.value.ident == "radar_b200:pc:indexOutOfRange"


def test_mex_fun_MTD_produce_rows():
    """fun_MTD_produce_rows(echo, lo, hi) == fun_MTD_produce(echo)(lo:hi, :) (MP/main_produce_dataset_win_xzr.m:37-40)."""
    echo = synth.s2_frame()
    want = mcode.fun_MTD_produce_mp(echo)
    got = Mex("fun_MTD_produce_rows")(echo, 3.0, 6.0)
    assert got.shape == (4, echo.shape[1])
    assert np.max(np.abs(got - want[2:6, :])) <= 1e-4 * np.max(np.abs(want))
    with pytest.raises(MexError) as e:
        Mex("fun_MTD_produce_rows")(echo, 3.0, float(echo.shape[0] + 1))
    assert e.value.ident == "radar_b200:mtdproduce:indexOutOfRange"


def test_mex_executeCFAR_and_cfar1d_bit_identical():
    rng = np.random.default_rng(5)
    x = rng.rayleigh(1.0, size=(64, 200))
    for _ in range(8):
        x[rng.integers(0, 64), rng.integers(0, 200)] = 30.0
    args = (5, 7, 3.0, 0, 5, 7, 3.0, 0, 1, 1)
    f, fv = Mex("executeCFAR")(x, *args, nargout=2)
    fo, fvo = mcode.executeCFAR(x, *args)
    assert np.array_equal(f, fo) and np.array_equal(fv, fvo)
    f1 = Mex("executeCFAR")(x, *args)                                 # single output
    assert np.array_equal(f1, fo)
    with pytest.raises(MexError) as e:
        Mex("executeCFAR")(x[:20], *args)
    assert e.value.ident == "radar_b200:cfar:indexOutOfRange"
    d = rng.rayleigh(1.0, size=(6, 60))
    assert np.array_equal(Mex("Function_CFAR1D_sub")(d, 5, 7, 1.5, 1), mcode.Function_CFAR1D_sub(d, 5, 7, 1.5, 1))
    assert np.array_equal(Mex("Function_CFAR1D_sub_fixCells")(d, 5, 7, 1.5, 0, [1, 4], [2, 30, 59]),
                          mcode.Function_CFAR1D_sub_fixCells(d, 5, 7, 1.5, 0, [1, 4], [2, 30, 59]))


def test_mex_motionParaMeasure():
    rng = np.random.default_rng(6)
    V, R, n0 = 48, 90, 2
    s = rng.rayleigh(1.0, size=(V, R)) + 0.5
    d = rng.normal(size=(V, R))
    flags = np.zeros((V, R))
    for v, r in ((10, 5), (30, 44), (n0 + 1, 0), (V - n0 - 1, R - 1)):
        s[v, r] += 25
        flags[v, r] = 1
    rScale, vScale, kValues = 6.0 * np.arange(R), 0.3 * (24 - np.arange(V)), 10 + rng.random((11, 12))
    args = (2, rScale, 6.0, 8, vScale, 0.3, 4, kValues, 3, 3.0, 1, 0.0, 0.0, n0)
    rE, vE, eE = Mex("motionParaMeasure")(s, d, flags, *args, nargout=3)
    wr, wv, we = mcode.motionParaMeasure(s, d, flags, *args)
    assert rE.shape == (4, 1)
    assert np.allclose(rE[:, 0], wr, rtol=1e-11) and np.allclose(vE[:, 0], wv, rtol=1e-11, atol=1e-12) and np.allclose(eE[:, 0], we, rtol=1e-11)
    none = Mex("motionParaMeasure")(s, d, np.zeros((V, R)), *args)
    assert none.size == 0


def test_mex_DMX_frame_process():
    """The DMX script block (FIR short pulse, circular matched filter, zero-padded MTD, sum / difference) through its gateway."""
    rng = np.random.default_rng(11)
    P, n_range, n_short, fft_num, mtd_fft, n0 = 48, 566, 62, 512, 64, 2
    left = np.rint(50 * _rc(rng, P, n_range))
    right = np.rint(50 * _rc(rng, P, n_range))
    mf = mcode.dmx_match_filter(mcode.load_ref("refDDCDataMF1"))
    win = mcode.hamming(P)
    want = mcode.dmx_frame(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0)
    got = Mex("DMX_frame_process")(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win.reshape(-1, 1), mtd_fft, n0, nargout=4)
    assert len(got) == 4
    for g, w in zip(got, want):
        _close(g, w)
    with pytest.raises(MexError) as e:
        Mex("DMX_frame_process")(left, right[:, :500], n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0, nargout=4)
    assert e.value.ident == "radar_b200:dmx:dimensionMismatch"
    with pytest.raises(MexError) as e:
        Mex("DMX_frame_process")(left, right, n_short, mcode.FILTER_COEF_INT, mf, 500, win, mtd_fft, n0, nargout=4)
    assert e.value.ident == "radar_b200:dmx:unsupported"


def test_mex_gateways_share_one_context_and_track_the_resident_plan():
    """All gateway binaries of a session hold references to ONE library context (rb200_shared_context_acquire).  A plan set
    through one gateway is recognised -- or replaced -- by another through the context's plan tag: interleaving
    fun_MTD_produce (1-arg plan, 1031 columns) with fun_lss_pulse_compression on a different geometry and back again
    must keep giving the oracle's results."""
    import ctypes
    import radar_signal_process_b200 as rsp
    rng = np.random.default_rng(9)
    p2, p3 = mcode.load_pulse_literals()
    echo_a = np.rint(100 * _rc(rng, 8, 1031))
    echo_b = np.rint(100 * _rc(rng, 4, 1024))
    want_a = mcode.fun_MTD_produce_mp(echo_a)
    want_b = mcode.fun_lss_pulse_compression_mp(echo_b, None, p2, p3)
    for _ in range(2):
        _close(Mex("fun_MTD_produce")(echo_a), want_a)
        _close(Mex("fun_lss_pulse_compression")(echo_b, 0, mcode.pulse1_mp(), p2, p3), want_b)
        _close(Mex("fun_lss_pulse_compression")(echo_a, 0, mcode.pulse1_mp(), p2, p3), mcode.fun_lss_pulse_compression_mp(echo_a, None, p2, p3))
    # the library hands every caller of the shared context the same handle
    lib = rsp.load()
    h1, h2 = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.rb200_shared_context_acquire(ctypes.byref(h1), 0) == 0
    assert lib.rb200_shared_context_acquire(ctypes.byref(h2), 0) == 0
    assert h1.value == h2.value and h1.value
    assert lib.rb200_get_plan_tag(h1) != 0
    assert lib.rb200_shared_context_release(0) == 0 and lib.rb200_shared_context_release(0) == 0
