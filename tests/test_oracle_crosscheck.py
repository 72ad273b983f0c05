"""Independent cross-checks of the oracle's restatements of MATLAB built-ins.

The reference holds no golden *outputs* for the hot path (only kaiser_win.mat, see test_oracle_golden.py), so the oracle's
building blocks are additionally compared here with independent implementations of the same documented definitions
(scipy.signal / numpy) and with known-answer values of the published algorithms.  This does not pin parity to a MATLAB run
(the oracle header still says "parity unpinned"); it removes the risk that two of our own restatements share one mistake.
"""
import numpy as np
import pytest
import scipy.signal as ss

from oracle import mcode, vec


def test_mround_is_half_away_from_zero():
    # MATLAB documentation of round(): "rounds away from zero for ties"
    cases = {0.5: 1, 1.5: 2, 2.5: 3, -0.5: -1, -1.5: -2, -2.5: -3, 0.49999: 0, 32.0: 32, 64 / 150: 0, 1536 / 150: 10, 256 / 20: 13}
    for x, want in cases.items():
        assert mcode.mround(x) == want, x


def test_zero_velocity_rows_known_answers():
    # MP/fun_0v_pressing.m:4,6  round(P/2) -/+ round(P/150), 1-based inclusive
    assert mcode.zero_v_rows(1536, 150) == (758, 778)
    assert mcode.zero_v_rows(64, 150) == (32, 32)
    assert mcode.zero_v_rows(256, 150) == (126, 130)
    assert mcode.zero_v_rows(64, 20) == (29, 35)


@pytest.mark.parametrize("nb,nx", [(35, 300), (1, 17), (67, 67), (5, 3)])
def test_filter_matches_scipy_lfilter(nb, nx):
    rng = np.random.default_rng(nb * 1000 + nx)
    b = rng.standard_normal(nb)
    x = rng.standard_normal(nx) + 1j * rng.standard_normal(nx)
    np.testing.assert_allclose(mcode.matlab_filter_fir(b, x), ss.lfilter(b, [1.0], x), rtol=1e-12, atol=1e-12)


def test_vectorised_fir_rows_match_scipy():
    rng = np.random.default_rng(5)
    b = mcode.FILTER_COEF_INT / 2048.0
    x = rng.standard_normal((4, 200)) + 1j * rng.standard_normal((4, 200))
    got = vec.fir_rows(x, b)
    np.testing.assert_allclose(got, ss.lfilter(b, [1.0], x, axis=1), rtol=1e-12, atol=1e-12)


def test_group_delay_matches_scipy_and_linear_phase_theory():
    b = mcode.FILTER_COEF_INT
    w, gd = ss.group_delay((b, [1.0]), w=512, whole=False)
    assert mcode.mround(float(np.mean(gd))) == mcode.grpdelay_mean_round(b)
    assert mcode.grpdelay_mean_round(b) == (b.size - 1) // 2          # symmetric FIR: constant delay (N-1)/2 = 17
    rng = np.random.default_rng(1)
    b2 = rng.standard_normal(9) + 3.0 * (np.arange(9) == 2)           # non-symmetric
    w, gd2 = ss.group_delay((b2, [1.0]), w=512, whole=False)
    assert mcode.grpdelay_mean_round(b2) == mcode.mround(float(np.mean(gd2)))


@pytest.mark.parametrize("L,M", [(67, 300), (7, 3), (1, 9), (160, 707)])
def test_pulse_compression_is_full_linear_convolution_with_matched_filter(L, M):
    # MP/fun_pulse_compression.m:4,16-22: ifft(fft(echo,n).*fft(conj(fliplr(s0)),n)), n = L+M-1  ==  conv(echo, conj(flip(s0)))
    rng = np.random.default_rng(L + M)
    s0 = rng.standard_normal(L) + 1j * rng.standard_normal(L)
    x = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    want = np.convolve(x, np.conj(s0[::-1]))
    got = mcode.fun_pulse_compression(s0, x)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    # the same thing as a correlation: out[n + L - 1] = sum_k x[n+k] conj(s0[k])
    corr = ss.correlate(x, s0, mode="full")
    np.testing.assert_allclose(got, corr, rtol=1e-10, atol=1e-10 * np.abs(want).max())


def test_vectorised_matched_filter_rows_match_correlation():
    rng = np.random.default_rng(3)
    ref = mcode.load_ref("refDDCDataMF1")
    x = rng.standard_normal((3, 500)) + 1j * rng.standard_normal((3, 500))
    got = vec.mf_rows(x, ref)
    for i in range(3):
        full = ss.correlate(x[i], ref, mode="full")            # index n + L - 1  <->  lag n
        want = full[ref.size - 1:ref.size - 1 + got.shape[1]]
        np.testing.assert_allclose(got[i, :want.size], want, rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("P", [8, 64, 256, 155])
def test_process_mtd_matches_numpy_definition(P):
    rng = np.random.default_rng(P)
    R = 13
    x = rng.standard_normal((P, R)) + 1j * rng.standard_normal((P, R))
    want = np.abs(np.fft.fftshift(np.fft.fft(x * np.kaiser(P, 8.0)[:, None], axis=0), axes=0))
    np.testing.assert_allclose(mcode.fun_Process_MTD(x, R, P), want, rtol=1e-10, atol=1e-10)


def test_kaiser_matches_numpy_and_scipy():
    for n in (2, 8, 64, 155, 256, 1536):
        np.testing.assert_allclose(mcode.kaiser(n, 8.0), np.kaiser(n, 8.0), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(mcode.kaiser(n, 8.0), ss.windows.kaiser(n, 8.0, sym=True), rtol=1e-12, atol=1e-15)


def test_tone_lands_on_the_fftshifted_bin():
    # a slow-time tone at Doppler bin m must peak at row (m + P/2) mod P  (fftshift, MP/fun_Process_MTD.m:24)
    P, R = 64, 4
    for m in (0, 1, 17, 32, 63):
        x = np.exp(2j * np.pi * m * np.arange(P) / P)[:, None] * np.ones((1, R))
        out = mcode.fun_Process_MTD(x, R, P)
        assert int(np.argmax(out[:, 0])) == (m + P // 2) % P


def test_cfar_known_answer_isolated_target():
    # CW/Function_CFAR1D_sub.m: a single strong cell on a flat floor is the only GO-CFAR hit; its neighbours inside
    # the guard band do not see it in their reference cells, cells whose reference window contains it are raised.
    x = np.ones((1, 60))
    x[0, 30] = 100.0
    f = mcode.Function_CFAR1D_sub(x, 5, 7, 5.0, 0)
    assert f[0, 30] == 1 and f.sum() == 1
    # threshold is inclusive (>=): a floor equal to T * mean flags everything
    g = mcode.Function_CFAR1D_sub(np.ones((1, 40)), 5, 7, 1.0, 0)
    assert g.sum() == 40


def test_cfar_edge_cells_use_the_other_side():
    # left edge: the leading window does not fit, the lagging mean is used for both (Function_CFAR1D_sub.m:24-40)
    x = np.ones((1, 50))
    x[0, 2] = 50.0
    x[0, 47] = 50.0
    f = mcode.Function_CFAR1D_sub(x, 5, 7, 5.0, 0)
    assert f[0, 2] == 1 and f[0, 47] == 1 and f.sum() == 2


def test_spline_matches_scipy_not_a_knot_and_reproduces_cubics():
    x = np.arange(7, dtype=float)
    y = 0.3 * x ** 3 - 2.0 * x ** 2 + x - 4.0                   # not-a-knot splines reproduce cubics exactly
    xq = np.linspace(0, 6, 41)
    np.testing.assert_allclose(mcode.matlab_spline_eval(y, xq), 0.3 * xq ** 3 - 2.0 * xq ** 2 + xq - 4.0, rtol=1e-10, atol=1e-10)
    y3 = np.array([1.0, 4.0, 2.0])                              # n = 3: the parabola through the three points
    c = np.polyfit(np.arange(3), y3, 2)
    np.testing.assert_allclose(mcode.matlab_spline_eval(y3, xq[:14]), np.polyval(c, xq[:14]), rtol=1e-10, atol=1e-10)


def test_unpack_ddc_known_answer():
    # FrameDataRead_xzr.m:122-129: int16 little-endian, I then Q, channel-fastest
    raw = np.array([1, -2, 3, -4, 5, -6, 7, -8], dtype="<i2").view(np.uint8)
    out = mcode.unpack_ddc_i16(raw, 2, 2)
    want = np.array([[1 - 2j, 3 - 4j], [5 - 6j, 7 - 8j]])
    np.testing.assert_array_equal(out, want)


def test_hamming_matches_scipy():
    import scipy.signal.windows as ssw
    for n in (2, 67, 96, 1536):
        np.testing.assert_allclose(mcode.hamming(n), ssw.hamming(n, sym=True), rtol=1e-13, atol=1e-15)


def test_dmx_long_pulse_is_circular_correlation_mod_fft_num():
    # CW/DMX_SignalProcessing_main_xzr.m:202,348-353: ifft(fft(x,512).*conj(fft(h,512))) = sum_k x[(n+k) mod 512] conj(h[k])
    rng = np.random.default_rng(4)
    P, n_long, N = 3, 504, 512
    h = mcode.dmx_match_filter(mcode.load_ref("refDDCDataMF1"))
    x = rng.standard_normal((P, n_long)) + 1j * rng.standard_normal((P, n_long))
    z = np.zeros((P, 0))
    _, _, s_long, d_long = mcode.dmx_frame(x, 2 * x, 0, mcode.FILTER_COEF_INT, h, N, np.ones(P), P, None)
    xp = np.concatenate([x, np.zeros((P, N - n_long))], axis=1)
    want = np.zeros((P, N), dtype=complex)
    for n in range(N):
        for k in range(h.size):
            want[:, n] += xp[:, (n + k) % N] * np.conj(h[k])
    mag = np.abs(np.fft.fft(want, axis=0))
    np.testing.assert_allclose(s_long, 3 * mag, rtol=1e-9, atol=1e-9)       # |L| + |R| with R = 2 L
    np.testing.assert_allclose(d_long, mag, rtol=1e-9, atol=1e-9)          # |R| - |L|
    assert np.isclose(np.sum(np.abs(h / mcode.kaiser(h.size, 4.5)) ** 2), 1.0)   # energy-normalised before the taper


def test_dmx_blanking_rows_and_short_pulse_fir():
    rng = np.random.default_rng(9)
    P, n_short, n_long, N, M, n0 = 8, 20, 30, 256, 16, 2
    x = rng.standard_normal((P, n_short + n_long)) + 1j * rng.standard_normal((P, n_short + n_long))
    h = np.ones(4)
    ss, ds, sl, dl = mcode.dmx_frame(x, x, n_short, mcode.FILTER_COEF_INT, h, N, mcode.hamming(P), M, n0)
    assert ss.shape == (M, n_short) and sl.shape == (M, N)
    assert np.all(ss[[0, 1, 2, 14, 15]] == 0) and np.all(sl[[0, 1, 2, 14, 15]] == 0) and np.all(ss[3:14] > 0)
    assert np.all(ds == 0) and np.all(dl == 0)                                # identical beams: zero difference channel
    want = np.abs(np.fft.fft(ss_ref(x[:, :n_short], P) * mcode.hamming(P)[:, None], M, axis=0)) * 2
    np.testing.assert_allclose(ss[3:14], want[3:14], rtol=1e-10, atol=1e-10)


def ss_ref(short, P):
    import scipy.signal as sig
    return sig.lfilter(mcode.FILTER_COEF_INT, [1.0], short, axis=1)           # un-normalised taps, no delay compensation (:344)
