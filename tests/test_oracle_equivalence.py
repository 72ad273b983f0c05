"""Loop-faithful transcription (oracle.mcode) == vectorised twin (oracle.vec), plus analytic known answers."""
import numpy as np
import pytest

from oracle import mcode, vec, synth


def _rand_c(rng, *shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


def test_unpack_matches_per_prt_rule():
    rng = np.random.default_rng(0)
    P, R, C = 3, 37, 16
    raw = rng.integers(-32768, 32768, size=(1, P, R, C, 2), dtype=np.int16)
    z = vec.unpack_wire(raw, 1, P, R, C)
    for p in range(P):
        sig, pad = mcode.ddc_payload_size(R, C)
        payload = raw[0, p].tobytes() + bytes(pad)
        one = mcode.unpack_ddc_i16(np.frombuffer(payload, dtype=np.uint8), R, C)
        assert np.array_equal(one, z[0, :, p, :].T)
    assert mcode.ddc_payload_size(4096, 16) == (262144, 0)
    assert mcode.ddc_payload_size(37, 16) == (2368, 0)
    assert mcode.ddc_payload_size(3, 1) == (12, 52)


def test_pc_closed_form_equals_fft_recipe():
    rng = np.random.default_rng(1)
    for L, M in ((75, 242), (160, 707), (67, 300), (1, 5), (7, 3)):
        s0 = _rand_c(rng, L)
        x = _rand_c(rng, M)
        full = mcode.fun_pulse_compression(s0, x)
        assert full.shape == (L + M - 1,)
        direct = np.array([sum(x[n + k] * np.conj(s0[k]) for k in range(L) if n + k < M) for n in range(M)])
        assert np.allclose(full[L - 1:L - 1 + M], direct, atol=1e-10)
        assert np.allclose(vec.mf_rows(x[None, :], s0)[0], direct, atol=1e-10)


def test_pc_impulse_gives_conj_reversed_reference():
    ref = mcode.load_ref("refDDCDataMF1")
    x = np.zeros(200, dtype=complex)
    x[120] = 1.0
    y = mcode.fun_pulse_compression(ref, x)
    # full convolution with conj(flip(ref)) delayed by 120
    assert np.allclose(y[120:120 + 67], np.conj(ref[::-1]), atol=1e-9)


def test_pc_peak_at_leading_edge():
    ref = mcode.load_ref("refDDCDataMF1")
    x = np.zeros(512, dtype=complex)
    x[100:167] = ref
    out = vec.single_pc(x[None, :], ref)[0]
    assert np.argmax(np.abs(out)) == 100
    assert np.isclose(out[100], np.sum(np.abs(ref) ** 2))


def test_lss_mp_loop_vs_vec():
    rng = np.random.default_rng(2)
    p2, p3 = mcode.load_pulse_literals()
    for n in (1031, 400):
        echo = _rand_c(rng, 5, n)
        a = mcode.fun_lss_pulse_compression_mp(echo, mcode.pulse1_mp(), p2, p3)
        b = vec.lss_pc_mp(echo, p2, p3)
        assert np.max(np.abs(a - b)) < 1e-11 * np.max(np.abs(a))
    with pytest.raises(mcode.MatlabError):
        mcode.fun_lss_pulse_compression_mp(echo, None, p2[:70], p3)
    with pytest.raises(mcode.MatlabError):
        vec.lss_pc_mp(echo, p2[:70], p3)


def test_lss_mtd_loop_vs_vec():
    rng = np.random.default_rng(3)
    params = dict(fs=25e6, B=20e6, tao=[0.16e-6, 8e-6, 28e-6], point_prt=[3404, 228, 723, 2453])
    pulse1, pulse2, pulse3 = mcode.ideal_pulses_mtd(params)
    assert pulse1.shape == (4,) and pulse2.shape == (200,) and pulse3.shape == (700,)   # MTD/fun_MTD_produce.m:67-69
    echo = _rand_c(rng, 3, 3404)
    a = mcode.fun_lss_pulse_compression_mtd(echo, pulse1, pulse2, pulse3, 228, 723, 2453)
    b = vec.lss_pc_mtd(echo, pulse2, pulse3, 228, 723, 2453)
    assert np.max(np.abs(a - b)) < 1e-11 * np.max(np.abs(a))
    # group-delay alignment: an impulse at column 10 of segment 1 peaks at column 10 (not 27)
    e = np.zeros((1, 3404), dtype=complex)
    e[0, 10] = 1
    o = vec.lss_pc_mtd(e, pulse2, pulse3, 228, 723, 2453)
    assert np.argmax(np.abs(o[0, :228])) == 10
    o5 = vec.lss_pc_mp(e[:, :1031], *mcode.load_pulse_literals())
    assert np.argmax(np.abs(o5[0, :82])) == 27            # 5-arg API leaves the FIR delayed by 17
    with pytest.raises(mcode.MatlabError):
        vec.lss_pc_mtd(echo, pulse2, pulse3, 228, 723, 2454)


def test_mtd_loop_vs_vec_and_tone():
    rng = np.random.default_rng(4)
    for P in (8, 64, 155, 1536):
        x = _rand_c(rng, P, 3)
        a = mcode.fun_Process_MTD(x, 3, P)
        b = vec.process_mtd(x, axis=0)
        assert np.max(np.abs(a - b)) < 1e-12 * np.max(np.abs(a))
    P = 64
    for k in (-32, -5, 0, 7, 31):
        x = np.exp(2j * np.pi * k * np.arange(P) / P)[:, None]
        m = mcode.fun_Process_MTD(x, 1, P)[:, 0]
        row = k + P // 2                                  # 0-based row of signed bin k
        assert np.argmax(m) == row
        assert np.isclose(m[row], mcode.kaiser(P, 8).sum())


def test_zero_v_variants():
    m = np.ones((64, 3))
    z = mcode.fun_0v_pressing(m, 150)
    assert np.array_equal(np.nonzero(z[:, 0] == 0)[0], [31])        # row 32 (1-based) = bin -1, DC untouched
    z = mcode.fun_0v_pressing(np.ones((155, 2)), 20)
    assert np.array_equal(np.nonzero(z[:, 0] == 0)[0], np.arange(69, 86))
    assert np.array_equal(vec.zero_v(m[None], 150)[0], mcode.fun_0v_pressing(m, 150))


def test_mti_and_istc():
    rng = np.random.default_rng(5)
    x = _rand_c(rng, 64, 9)
    a = mcode.fun_Process_MTI(x)
    assert np.array_equal(a, vec.mti(x, 30, axis=0))
    assert np.all(a[34:] == 0) and np.array_equal(a[0], x[30] - x[0])
    stc, y = mcode.fun_iSTC(x, [6.0, 3.0])
    assert np.allclose(y[:, 0], x[:, 0] * 10 ** 0.3) and np.array_equal(y[:, 2:], x[:, 2:])
    assert np.allclose(vec.istc(x, [6.0, 3.0]), y)


@pytest.mark.parametrize("seed", range(12))
def test_cfar_loop_vs_dense(seed):
    rng = np.random.default_rng(100 + seed)
    V, R = int(rng.integers(26, 40)), int(rng.integers(24, 60))
    n0 = int(rng.integers(0, 2))
    x = rng.rayleigh(1.0, size=(V, R))
    # strong peaks, exact ties and zero rows to exercise >=, first-max and 0>=0
    for _ in range(6):
        v, r = int(rng.integers(0, V)), int(rng.integers(0, R))
        x[v, r] = 40.0
        if r + 1 < R and rng.random() < 0.5:
            x[v, r + 1] = 40.0
    x[V // 2, :] = 0.0
    meth = int(seed % 2)
    T = 3.0 if seed % 3 else 1.0
    refV, gV = (5, 7) if V - 2 * n0 - 1 >= 24 else (2, 1)
    args = (5, 7, T, meth, refV, gV, T, meth, n0, 1)
    f1, v1 = mcode.executeCFAR(x, *args)
    f2, v2 = vec.execute_cfar(x, *args)
    assert np.array_equal(v1, v2)
    assert np.array_equal(f1, f2)
    f3, v3 = mcode.executeCFAR(x, *args[:-1], 0)
    assert np.array_equal(f3, v3) and np.array_equal(v3, v1)
    f4, _ = vec.execute_cfar(x, *args[:-1], 0)
    assert np.array_equal(f4, f3)


def test_cfar1d_vs_loop_and_edges():
    rng = np.random.default_rng(7)
    d = rng.rayleigh(1.0, size=(4, 50))
    for meth in (0, 1):
        a = mcode.Function_CFAR1D_sub(d, 5, 7, 1.2, meth)
        b = vec.cfar1d_last(d, 5, 7, 1.2, meth)
        assert np.array_equal(a, b)
    # constant background: flagged iff T <= 1
    c = np.full((2, 30), 3.0)
    assert mcode.Function_CFAR1D_sub(c, 5, 7, 1.0, 0).all()
    assert not mcode.Function_CFAR1D_sub(c, 5, 7, 1.0000001, 0).any()
    # all-zero neighbourhood and zero CUT is flagged (0 >= 0), SURVEY 7.4-4
    assert mcode.Function_CFAR1D_sub(np.zeros((1, 24)), 5, 7, 5.0, 0).all()
    # axis shorter than 2*(ref+guard) raises like MATLAB's index error
    with pytest.raises(mcode.MatlabError):
        mcode.Function_CFAR1D_sub(np.zeros((1, 23)), 5, 7, 5.0, 0)
    with pytest.raises(mcode.MatlabError):
        vec.cfar1d_last(np.zeros((1, 23)), 5, 7, 5.0, 0)
    # edge substitution: spike at column 3 uses the right window only
    e = np.ones((1, 40))
    e[0, 2] = 10.0
    assert mcode.Function_CFAR1D_sub(e, 5, 7, 5.0, 0)[0, 2] == 1
    # fixCells only touches the listed cells
    f = mcode.Function_CFAR1D_sub_fixCells(e, 5, 7, 5.0, 0, 1, [2, 3, 4])
    assert f.sum() == 1 and f[0, 2] == 1


def test_isolated_spike_single_2d_detection():
    x = np.ones((64, 128))
    x[20, 60] = 100.0
    f, fv = vec.execute_cfar(x, 5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)
    assert f.sum() == 1 and f[20, 60] == 1
    assert fv.sum() == 1


def test_fun_cfarflag_segments():
    rng = np.random.default_rng(9)
    x = rng.rayleigh(1.0, size=(60, 1031))
    x[30, 100] = 50
    x[31, 500] = 50
    args = (5, 7, 5.0, 0, 5, 7, 5.0, 0, 2, 1)
    a = mcode.fun_CFARflag(x, *args)
    b = vec.cfar_flag_segments(x, args)
    assert np.array_equal(a, b)
    assert a[:, 868:].sum() == 0 and a[30, 100] == 1 and a[31, 500] == 1


def test_full_chain_mp_small():
    rng = np.random.default_rng(11)
    echo = np.rint(rng.normal(0, 100, size=(16, 1031))) + 1j * np.rint(rng.normal(0, 100, size=(16, 1031)))
    a = mcode.fun_MTD_produce_mp(echo)
    p2, p3 = mcode.load_pulse_literals()
    b = vec.zero_v(vec.process_mtd(vec.lss_pc_mp(echo, p2, p3), axis=0), 150, axis=0)
    assert a.shape == (16, 1031)
    assert np.max(np.abs(a - b)) < 1e-11 * np.max(a)


def test_synth_s3_wire_and_chain_tiny():
    raw, tg = synth.s3_batch(1, P=64, R=512, C=2, n_targets=2, r_lo=50, r_hi=400)
    assert raw.shape == (1, 64, 512, 2, 2) and raw.dtype == np.int16
    ref = mcode.load_ref("refDDCDataMF1")
    out = vec.chain(raw, 1, 64, 512, 2, ("single", ref), synth.cfar_tuple(synth.S3_CFAR), near_tol=1e-4)
    assert out["rdm"].shape == (1, 2, 64, 512)
    # every injected target is detected at its leading-edge range cell and Doppler row
    for lane, r0, k, snr in tg[0]:
        assert out["flag"][0, lane, k + 32, max(r0 - 1, 0):r0 + 2].sum() >= 1, (lane, r0, k, snr)
    assert np.all(out["rdm"][0, :, 31, :] == 0)


def test_dbf_weighting_matches_per_prt_matrix_product():
    # FrameDataRead_xzr.m:156-158: current_sig_data_DBF = sig_data_C * DBF_coeffs_data_C.'  (n x 16) * (16 x 13)
    rng = np.random.default_rng(21)
    P, R, C, NB = 3, 50, 16, 13
    raw = rng.integers(-2000, 2000, size=(1, P, R, C, 2), dtype=np.int16)
    W = _rand_c(rng, NB, C)
    x = vec.unpack_wire(raw, 1, P, R, C)
    beams = vec.dbf_weighting(x, W)
    assert beams.shape == (1, NB, P, R)
    for p in range(P):
        sig_c = mcode.unpack_ddc_i16(np.frombuffer(raw[0, p].tobytes(), dtype=np.uint8), R, C)      # n x 16
        want = sig_c @ W.T                                                                         # non-conjugate transpose
        assert np.allclose(beams[0, :, p, :].T, want, rtol=0, atol=1e-9)


def test_unpack_dbf24_known_words():
    # FrameDataRead_xzr.m:111-119,130-135: 13 beams -> 78 data bytes + 2 pad bytes per sample; > 2^23 goes negative
    assert mcode.dbf24_payload_size(3404, 13) == (3404 * 80, (64 - (3404 * 80) % 64) % 64, 2)
    words = {0x000001: 1, 0x7FFFFF: 8388607, 0x800000: 8388608, 0x800001: -8388607, 0xFFFFFF: -1, 0x000000: 0}
    ch, n = 13, 2
    sig, pad, osp = mcode.dbf24_payload_size(n, ch)
    payload = np.zeros(sig + pad, dtype=np.uint8)
    keys = list(words)
    for r in range(n):
        for v in range(2 * ch):
            w = keys[(r * 7 + v) % len(keys)]
            o = r * (6 * ch + osp) + 3 * v
            payload[o:o + 3] = (w & 255, (w >> 8) & 255, (w >> 16) & 255)
    z = mcode.unpack_dbf24(payload, n, ch)
    assert z.shape == (n, ch)
    for r in range(n):
        for b in range(ch):
            assert z[r, b].real == words[keys[(r * 7 + 2 * b) % len(keys)]]
            assert z[r, b].imag == words[keys[(r * 7 + 2 * b + 1) % len(keys)]]
    # aligned case keeps the 8-byte per-sample padding and yields one extra (zero) column, as the M slicing does
    assert mcode.dbf24_payload_size(4, 4)[2] == 8
    z4 = mcode.unpack_dbf24(np.zeros(sum(mcode.dbf24_payload_size(4, 4)[:2]), dtype=np.uint8), 4, 4)
    assert z4.shape == (4, 5)


def test_motion_para_measure_oracle_known_answers():
    # a separable Gaussian bump centred between cells: the spline arg-max lands on the nearest 1/times grid point
    V, R = 40, 60
    v0, r0 = 17.3, 31.6                      # 1-based fractional peak position
    vv, rr = np.meshgrid(np.arange(1, V + 1), np.arange(1, R + 1), indexing="ij")
    s = 100 * np.exp(-((vv - v0) ** 2) / 8.0 - ((rr - r0) ** 2) / 8.0)
    d = 0.25 * s
    flags = np.zeros((V, R))
    flags[16, 31] = 1                        # cell (17, 32)
    rScale = 6.0 * np.arange(R)
    vScale = 0.5 * (20 - np.arange(V))
    kValues = np.arange(1, 23, dtype=float).reshape(11, 2, order="F")
    rE, vE, eE = mcode.motionParaMeasure(s, d, flags, 2, rScale, 6.0, 8, vScale, 0.5, 4, kValues, 1, 3.0, 2, 0.1, 0.2, 3)
    assert rE.shape == (1,)
    assert abs(rE[0] - (rScale[31] + (31.625 - 32) * 6.0)) < 1e-9          # nearest 1/8 to 31.6
    assert abs(vE[0] - (vScale[16] - 0.25 * 0.5)) < 1e-9                   # nearest 1/4 to 17.3 is 17.25
    assert abs(eE[0] - (1 * 3.0 + 2.5 - 0.25 * kValues[2, 1] + 0.1 + 0.2)) < 1e-12
    # find() order is column-major and edges clamp the interpolation window
    flags[:] = 0
    flags[[4, 38, 10], [0, 59, 30]] = 1
    rE, vE, eE = mcode.motionParaMeasure(s + 1, d, flags, 2, rScale, 6.0, 8, vScale, 0.5, 4, kValues, 0, 3.0, 0, 0, 0, 3)
    assert rE.shape == (3,) and np.all(np.isfinite(rE)) and np.all(np.isfinite(vE))
    # spline restatement: n = 5 not-a-knot reproduces a cubic exactly
    x = np.arange(5.0)
    y = 2 - x + 0.5 * x ** 2 - 0.1 * x ** 3
    xq = np.linspace(0, 4, 33)
    assert np.allclose(mcode.matlab_spline_eval(y, xq), 2 - xq + 0.5 * xq ** 2 - 0.1 * xq ** 3, atol=1e-12)
