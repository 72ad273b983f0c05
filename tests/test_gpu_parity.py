"""GPU parity: the CUDA path, called through the C ABI (ctypes host mirror), against the CPU oracle.

Tolerances (BASELINE.json north_star):
  * int16 I/Q unpack: bit-exact
  * pulse-compressed samples / RDM magnitudes: max|gpu - ref| <= 1e-4 * max|ref|   (fp32 vs double)
  * CFAR flags: identical, except cells whose decision lies within 1e-4 (relative) of a threshold,
    which are counted and printed.  The MATLAB-layout double entry points must be bit-identical.
"""
import numpy as np
import pytest

from oracle import mcode, synth, vec

pytestmark = pytest.mark.gpu

RTOL = 1e-4
NEAR_FRAC = 1e-3        # differing flags (all of them excused by a near-threshold mask) per cell, upper bound
NEAR_MASK_FRAC = 0.25   # cells the oracle may mark "near a threshold or a tie" (zeroed 0-v rows tie exactly and are dilated over the
                        # 2-D election: 1.7 % of S3, 11.5 % of the S1 crop); what is bounded tightly is the number of cells that DIFFER


def _rand_c(rng, *shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


def _close(a, b, tol=RTOL):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.max(np.abs(b))
    err = np.max(np.abs(a - b))
    assert err <= tol * max(scale, 1e-300), "max err %.3e vs scale %.3e (rel %.3e)" % (err, scale, err / max(scale, 1e-300))
    return err / max(scale, 1e-300)


# ---------------------------------------------------------------------------------------------------
def test_library_is_native_and_loaded(lib):
    import os
    assert os.path.exists(lib.LIB_PATH)
    assert lib.load().rb200_version() >= 1
    ctx = lib.default_context()
    assert ctx._h.value


def test_unpack_bit_exact(lib):
    rng = np.random.default_rng(0)
    P, R, C, B = 4, 333, 16, 2
    raw = rng.integers(-32768, 32768, size=(B, P, R, C, 2), dtype=np.int16)
    raw[0, 0, 0, 0] = (-32768, 32767)
    with lib.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B) as ctx:
        got = ctx.unpack(raw, B)
    want = vec.unpack_wire(raw, B, P, R, C)
    assert np.array_equal(got.astype(np.complex128), want)


@pytest.mark.parametrize("L,M", [(67, 300), (75, 242), (160, 707), (7, 3), (1, 5), (200, 723), (700, 2453), (35, 82), (5000, 300)])
def test_fun_pulse_compression(lib, L, M):
    rng = np.random.default_rng(L * 1000 + M)
    s0 = _rand_c(rng, L)
    x = _rand_c(rng, M)
    got = lib.fun_pulse_compression(s0, x)
    want = mcode.fun_pulse_compression(s0, x)
    assert got.shape == (1, L + M - 1)
    _close(got[0], want)


def test_fun_pulse_compression_real_inputs_and_reference_chirp(lib):
    ref = mcode.load_ref("refDDCDataMF1")
    x = np.zeros(512)
    x[100] = 3.0
    got = lib.fun_pulse_compression(ref, x)[0]
    want = mcode.fun_pulse_compression(ref, x)
    _close(got, want)
    assert np.argmax(np.abs(got)) in range(100, 100 + 67)


def test_fun_lss_pulse_compression_5arg(lib):
    rng = np.random.default_rng(5)
    p2, p3 = mcode.load_pulse_literals()
    for n in (1031, 1024, 400):
        echo = np.rint(200 * _rand_c(rng, 6, n))
        got = lib.fun_lss_pulse_compression(echo, 0, mcode.pulse1_mp(), p2, p3)
        want = mcode.fun_lss_pulse_compression_mp(echo, None, p2, p3)
        _close(got, want)
    with pytest.raises(lib.MatlabDimensionError):
        lib.fun_lss_pulse_compression(echo, 0, None, p2[:70], p3)
    with pytest.raises(lib.MatlabIndexError):
        lib.fun_lss_pulse_compression(echo[:, :300], 0, None, p2, p3)


def test_fun_lss_pulse_compression_9arg(lib):
    rng = np.random.default_rng(6)
    params = dict(fs=25e6, B=20e6, tao=[0.16e-6, 8e-6, 28e-6], point_prt=[3404, 228, 723, 2453])
    pulse1, pulse2, pulse3 = mcode.ideal_pulses_mtd(params)
    echo = np.rint(100 * _rand_c(rng, 4, 3404))
    got = lib.fun_lss_pulse_compression(echo, params, 0, pulse1, pulse2, pulse3, 228, 723, 2453)
    want = mcode.fun_lss_pulse_compression_mtd(echo, pulse1, pulse2, pulse3, 228, 723, 2453)
    _close(got, want)
    # shorter third segment leaves the tail zero
    got = lib.fun_lss_pulse_compression(echo, params, 0, pulse1, pulse2, pulse3, 228, 723, 2000)
    want = mcode.fun_lss_pulse_compression_mtd(echo, pulse1, pulse2, pulse3, 228, 723, 2000)
    _close(got, want)
    assert np.all(got[:, 228 + 723 + 2000:] == 0)
    with pytest.raises(lib.MatlabIndexError):
        lib.fun_lss_pulse_compression(echo, params, 0, pulse1, pulse2, pulse3, 228, 723, 2454)


@pytest.mark.parametrize("P", [8, 64, 256, 155, 1536, 7, 1])
def test_fun_Process_MTD(lib, P):
    rng = np.random.default_rng(P)
    cols = 37
    x = 1000 * _rand_c(rng, P, cols)
    got = lib.fun_Process_MTD(x, cols, P)
    want = mcode.fun_Process_MTD(x, cols, P)
    _close(got, want)
    if P > 1:
        got = lib.fun_Process_MTD(x, 5, P)          # Len_PRT < columns
        _close(got, want[:, :5])
    with pytest.raises(lib.MatlabIndexError):
        lib.fun_Process_MTD(x, cols + 1, P)
    if P > 2:
        with pytest.raises(lib.MatlabDimensionError):
            lib.fun_Process_MTD(x, cols, P - 1)


def test_fun_Process_MTD_tone_lands_on_shifted_bin(lib):
    P = 64
    for k in (-32, -1, 0, 5, 31):
        x = np.exp(2j * np.pi * k * np.arange(P) / P)[:, None] * np.ones((1, 3))
        m = lib.fun_Process_MTD(x, 3, P)
        assert np.argmax(m[:, 1]) == k + 32
        assert abs(m[k + 32, 1] - mcode.kaiser(P, 8).sum()) < 1e-4 * P


@pytest.mark.parametrize("P,div", [(1536, 150), (64, 150), (8, 150), (256, 150), (155, 20)])
def test_fun_0v_pressing(lib, P, div):
    rng = np.random.default_rng(P)
    m = rng.rayleigh(1.0, size=(P, 9))
    got = lib.fun_0v_pressing(m, div)
    assert np.array_equal(got, mcode.fun_0v_pressing(m, div))


def test_fun_MTD_produce_rows_S1(lib):
    """Crop-aware entry (rb200_mtd_produce_rows_z): rows 691:845 of the 1536 x 1031 frame (MP/main_produce_dataset_win_xzr.m:
    37-40) with the slow-time transform first and the pulse compression on the kept rows only, against the oracle's full
    fun_MTD_produce cropped afterwards and against the uncropped GPU call; the MATLAB index error for a crop outside 1..P."""
    echo = synth.s1_frame(0)
    p2, p3 = mcode.load_pulse_literals()
    want = vec.zero_v(vec.process_mtd(vec.lss_pc_mp(echo, p2, p3), axis=0), 150, axis=0)
    got = lib.fun_MTD_produce_rows(echo, 691, 845)
    assert got.shape == (155, echo.shape[1])
    scale = np.max(np.abs(want))
    err = np.max(np.abs(got - want[690:845, :])) / scale
    print("S1 crop-aware RDM rel err %.2e (full-frame scale)" % err)
    assert err <= RTOL
    assert np.all(got[757 - 690:778 - 690, :] == 0)                  # zero-velocity rows 758:778 fall inside the crop
    full = lib.fun_MTD_produce(echo)
    assert np.max(np.abs(got - full[690:845, :])) <= 1e-5 * scale    # same result whichever operator runs first
    one = lib.fun_MTD_produce_rows(echo, 1, 1)
    assert one.shape == (1, echo.shape[1]) and np.max(np.abs(one - want[0:1, :])) <= RTOL * scale
    with pytest.raises(lib.MatlabIndexError):
        lib.fun_MTD_produce_rows(echo, 691, 1537)
    with pytest.raises(lib.MatlabIndexError):
        lib.fun_MTD_produce_rows(echo, 0, 10)


def test_fun_MTD_produce_S1(lib):
    """Config 1 stand-in: 1536 x 1031 frame through the 1-arg API (literal pulses, segments 82/242/707)."""
    echo = synth.s1_frame(0)
    got = lib.fun_MTD_produce(echo)
    p2, p3 = mcode.load_pulse_literals()
    want = vec.zero_v(vec.process_mtd(vec.lss_pc_mp(echo, p2, p3), axis=0), 150, axis=0)
    rel = _close(got, want)
    print("S1 RDM rel err %.2e" % rel)
    assert np.all(got[757:778, :] == 0)
    # then main_produce-style crop, CW/ 0-v, fun_CFARflag (CW/main_cfar.m:86-93,142-161)
    crop_g = lib.fun_0v_pressing(np.abs(got[690:845, :]), 20)
    crop_w = vec.zero_v(want[690:845, :], 20, axis=0)
    args = synth.cfar_tuple(synth.S1_CFAR)
    fg = np.zeros(crop_g.shape)
    for a, b in ((1, 82), (83, 318), (319, 868)):
        f, _ = lib.executeCFAR(crop_g[:, a - 1:b], *args)
        fg[:, a - 1:b] = f
    # same input (the GPU RDM) through the oracle must give identical flags: double path is exact
    assert np.array_equal(fg, vec.cfar_flag_segments(crop_g, args))
    # against the oracle's own RDM only near-threshold cells may differ (CW/main_cfar.m:86-93,142-161), and few of them
    want_f = _segmented_cfar_oracle(crop_w, args, ((0, 82), (82, 318), (318, 868)))
    diff = fg.astype(bool) != want_f["flag"]
    unexcused = diff & ~want_f["near"]
    print("S1 flags: gpu %d oracle %d differing %d unexcused %d near-mask %d" %
          (fg.sum(), want_f["flag"].sum(), diff.sum(), unexcused.sum(), want_f["near"].sum()))
    assert np.array_equal(want_f["flag"], vec.cfar_flag_segments(crop_w, args).astype(bool))
    assert unexcused.sum() == 0
    assert diff.sum() <= NEAR_FRAC * fg.size and want_f["near"].sum() <= NEAR_MASK_FRAC * fg.size


def test_fun_MTD_produce_S2_and_cfar(lib):
    echo = synth.s2_frame()
    got = lib.fun_MTD_produce(echo)
    want = mcode.fun_MTD_produce_mp(echo)
    _close(got, want)
    assert np.all(got[3, :] == 0) and np.all(got[4, :] > 0)       # row 4 (1-based) zeroed, DC untouched
    args = synth.cfar_tuple(synth.S2_CFAR)
    f, fv = lib.executeCFAR(got, *args)
    fo, fvo = mcode.executeCFAR(got, *args)
    assert np.array_equal(f, fo) and np.array_equal(fv, fvo)


def test_fun_MTD_produce_2arg(lib):
    rng = np.random.default_rng(12)
    params = dict(prt=1e-4, prf=1e4, prtNum=32, fs=25e6, deltaR=6.0, fc=9.4e9, wavelength=0.0319, B=20e6,
                  tao=[0.16e-6, 8e-6, 28e-6], point_prt=[3404, 228, 723, 2453], debug=dict(show_PC=0, show_FFT=0, graph=0))
    echo = np.rint(100 * _rand_c(rng, 32, 3404))
    got = lib.fun_MTD_produce(echo, params)
    want = mcode.fun_MTD_produce_mtd(echo, params)
    _close(got, want)


@pytest.mark.parametrize("seed", range(6))
def test_executeCFAR_bit_identical(lib, seed):
    rng = np.random.default_rng(200 + seed)
    V, R = int(rng.integers(30, 70)), int(rng.integers(24, 300))
    n0 = int(rng.integers(0, 3))
    x = rng.rayleigh(1.0, size=(V, R))
    for _ in range(10):
        v, r = int(rng.integers(0, V)), int(rng.integers(0, R))
        x[v, r] = 40.0
        if r + 1 < R and rng.random() < 0.5:
            x[v, r + 1] = 40.0                      # exact tie -> first max
    x[V // 2, :] = 0.0                              # 0 >= 0 is flagged in the range stage
    meth = seed % 2
    T = 3.0 if seed % 3 else 1.0
    args = (5, 7, T, meth, 5, 7, T, meth, n0, 1)
    f, fv = lib.executeCFAR(x, *args)
    fo, fvo = mcode.executeCFAR(x, *args)
    assert np.array_equal(fv, fvo)
    assert np.array_equal(f, fo)
    f0, fv0 = lib.executeCFAR(x, *args[:-1], 0)
    assert np.array_equal(f0, fvo) and np.array_equal(fv0, fvo)


def test_executeCFAR_errors_and_edges(lib):
    x = np.ones((20, 40))
    with pytest.raises(lib.MatlabIndexError):       # velocity axis (19 rows) shorter than 24
        lib.executeCFAR(x, 5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)
    x = np.ones((64, 128))
    x[20, 60] = 100.0
    f, fv = lib.executeCFAR(x, 5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)
    assert f.sum() == 1 and f[20, 60] == 1 and fv.sum() == 1
    c = np.full((64, 64), 3.0)                      # constant background: flagged iff T <= 1
    f, fv = lib.executeCFAR(c, 5, 7, 1.0, 0, 5, 7, 1.0, 0, 0, 1)
    fo, fvo = mcode.executeCFAR(c, 5, 7, 1.0, 0, 5, 7, 1.0, 0, 0, 1)
    assert np.array_equal(f, fo) and np.array_equal(fv, fvo) and fv[1:].all()


def test_Function_CFAR1D_sub_and_fixCells(lib):
    rng = np.random.default_rng(7)
    d = rng.rayleigh(1.0, size=(9, 50))
    for meth in (0, 1):
        assert np.array_equal(lib.Function_CFAR1D_sub(d, 5, 7, 1.2, meth), mcode.Function_CFAR1D_sub(d, 5, 7, 1.2, meth))
    assert np.array_equal(lib.Function_CFAR1D_sub_fixCells(d, 5, 7, 1.2, 0, [2, 5], [1, 25, 50]),
                          mcode.Function_CFAR1D_sub_fixCells(d, 5, 7, 1.2, 0, [2, 5], [1, 25, 50]))
    assert lib.Function_CFAR1D_sub(np.zeros((1, 24)), 5, 7, 5.0, 0).all()
    with pytest.raises(lib.MatlabIndexError):
        lib.Function_CFAR1D_sub(np.zeros((1, 23)), 5, 7, 5.0, 0)
    with pytest.raises(lib.MatlabIndexError):
        lib.Function_CFAR1D_sub_fixCells(d, 5, 7, 1.2, 0, [10], [1])


# ---------------------------------------------------------------------------------------------------
# the batched wire-format chain (benchmark path)
# ---------------------------------------------------------------------------------------------------
def _chain_ctx(lib, P, R, C, B, plan, cfar, **kw):
    ctx = lib.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, **kw)
    ctx.set_waveform(plan)
    ctx.set_cfar(*cfar)
    return ctx


def _compare_flags(got_dets, out, B, C, V, R, lib):
    flag, flagv = lib.dets_to_flags(got_dets, B, C, V, R)
    dv = flagv != out["flagV"]
    d2 = flag != out["flag"]
    bad_v = dv & ~out["nearV"]
    bad_2 = d2 & ~out["near"]
    print("flags: V-stage gpu %d oracle %d differ %d (unexcused %d) | 2-D gpu %d oracle %d differ %d (unexcused %d)" %
          (flagv.sum(), out["flagV"].sum(), dv.sum(), bad_v.sum(), flag.sum(), out["flag"].sum(), d2.sum(), bad_2.sum()))
    assert bad_v.sum() == 0
    assert bad_2.sum() == 0
    # the excused set itself is bounded: cells that differ (all of them inside a near-threshold mask, asserted above) may be
    # at most 1e-3 of all cells -- a kernel that pushed 5 % of the cells across a threshold fails here even if excused
    cells = flag.size
    assert dv.sum() <= NEAR_FRAC * cells and d2.sum() <= NEAR_FRAC * cells, (int(dv.sum()), int(d2.sum()), cells)
    assert out["nearV"].sum() <= NEAR_MASK_FRAC * cells and out["near"].sum() <= NEAR_MASK_FRAC * cells, \
        (int(out["nearV"].sum()), int(out["near"].sum()), cells)
    return int(dv.sum()), int(d2.sum())


def test_chain_S3_two_cpis(lib):
    """Headline shape 64 x 4096 x 16, two CPIs, plan 'single' with refDDCDataMF1."""
    P, R, C, B = 64, 4096, 16, 2
    raw, targets = synth.s3_batch(B)
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("single", ref), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=B) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
        pc = ctx.debug_fetch_pc(B - 1)
        assert ctx.last_launch_count() >= 3 and ctx.last_device_ms() > 0
    print("S3 PC rel err %.2e" % _close(pc, out["pc"][B - 1]))
    print("S3 RDM rel err %.2e" % _close(rdm, out["rdm"]))
    assert n == len(dets) and n > 0
    _compare_flags(dets, out, B, C, P, R, lib)
    # every injected target is found within one range cell of its leading edge
    flag, _ = lib.dets_to_flags(dets, B, C, P, R)
    missed = sum(1 for b in range(B) for lane, r0, k, snr in targets[b]
                 if flag[b, lane, k + 32, max(r0 - 1, 0):r0 + 2].sum() == 0 and out["flag"][b, lane, k + 32, max(r0 - 1, 0):r0 + 2].sum() > 0)
    assert missed == 0


def test_chain_lss_plan_and_odd_lane_count(lib):
    """Secondary plan 'lss' (82/242/rest with FIR + literal pulses) on 13 lanes."""
    P, R, C, B = 64, 1031, 13, 1
    rng = np.random.default_rng(3)
    raw = rng.integers(-300, 300, size=(B, P, R, C, 2), dtype=np.int16)
    p2, p3 = mcode.load_pulse_literals()
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("lss_mp", p2, p3), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_mp(R, p2, p3), cfar) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
    _close(rdm, out["rdm"])
    _compare_flags(dets, out, B, C, P, R, lib)


def test_chain_S5_small_istc_mti(lib):
    """DBF-mode long CPI (P=256) with iSTC and MTI on, reduced range extent."""
    P, R, C, B = 256, 2048, 4, 1
    ref = mcode.load_ref("refDBFDataMF1")
    raw, _ = synth.s3_cpi(0, P=P, R=R, C=C, ref=ref, seed0=5000, r_lo=100, r_hi=R - 200, exclude=(-3, -2, -1, 0, 1, 2, 3))
    raw = raw[None]
    stc = synth.s5_stc_curve()
    cfar = synth.cfar_tuple(synth.S5_CFAR)
    out = vec.chain(raw, B, P, R, C, ("single", ref), cfar, stc=stc, mti_lag=30, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, mti_lag=30) as ctx:
        ctx.set_stc(stc)
        rdm, dets, n = ctx.chain(raw, B)
    _close(rdm, out["rdm"])
    assert np.all(rdm[:, :, 125:130, :] == 0)
    _compare_flags(dets, out, B, C, P, R, lib)


def test_chain_full_size_properties(lib):
    """BASELINE full-size batch (8 CPIs of 64 x 4096 x 16): size-independent properties."""
    P, R, C, B = 64, 4096, 16, 8
    raw, _ = synth.s3_batch(2)
    raw = np.concatenate([raw] * 4, axis=0)                     # cpi 0,1,0,1,...
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)

    def key(d):
        return np.sort(d, order=["cpi", "lane", "v", "r", "kind"])

    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=3) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
        # (a) replicated CPIs give replicated outputs (chunk boundaries fall mid-batch)
        assert np.array_equal(rdm[0], rdm[2]) and np.array_equal(rdm[1], rdm[7])
        d0 = key(dets[dets["cpi"] == 0])
        d6 = key(dets[dets["cpi"] == 6])
        assert len(d0) == len(d6) and all(np.array_equal(d0[f], d6[f]) for f in ("lane", "v", "r", "kind", "amp"))
        # (b) exact power-of-two linearity: 2x input -> 2x RDM, identical detections
        small = (raw[:2] // 4).astype(np.int16)
        r1, dd1, _ = ctx.chain(small, 2)
        r2, dd2, _ = ctx.chain((small * 2).astype(np.int16), 2)
        assert np.array_equal(r2, 2 * r1)
        k1, k2 = key(dd1), key(dd2)
        assert all(np.array_equal(k1[f], k2[f]) for f in ("cpi", "lane", "v", "r", "kind"))
    # (c) chunk size does not change results
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=8) as ctx:
        rdm8, dets8, _ = ctx.chain(raw, B)
    assert np.array_equal(rdm8, rdm)
    a, b = key(dets), key(dets8)
    assert len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)
    # (d) impulse response: a unit impulse on every PRT compresses to conj(ref) reversed before the impulse
    imp = np.zeros((1, P, R, C, 2), dtype=np.int16)
    imp[0, :, 1000, :, 0] = 1
    with _chain_ctx(lib, P, R, C, 1, lib.waveforms.segments_single(R, ref), cfar) as ctx:
        ctx.set_debug_keep_pc(True)
        ctx.chain(imp, 1, want_rdm=False, allow_overflow=True)   # all-zero background: 0 >= 0 flags everything
        pc = ctx.debug_fetch_pc(0)
    want = np.zeros(R, dtype=complex)
    want[1000 - 66:1001] = np.conj(ref[::-1])
    _close(pc[5, 17], want)


def test_chain_bench_configuration_64_cpis_spot_checked(lib):
    """The exact bench configuration (64 CPIs of 64 x 4096 x 16 per call, default chunk size, default slot streams, host
    buffers) with four randomly chosen CPIs compared cell by cell against the oracle (RDM and both flag matrices)."""
    P, R, C, B, D = 64, 4096, 16, 64, 8
    distinct, _ = synth.s3_batch(D)
    order = np.random.default_rng(64).permutation(B) % D            # every distinct CPI appears 8 times, shuffled
    raw = np.ascontiguousarray(distinct[order])
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, max_det=1 << 20) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
    assert n == len(dets) and n > 0
    picks = np.random.default_rng(65).choice(B, size=4, replace=False)
    assert len({int(order[i]) for i in picks}) >= 3
    for i in picks:
        src = int(order[i])
        out = vec.chain(distinct[src:src + 1], 1, P, R, C, ("single", ref), cfar, near_tol=RTOL)
        _close(rdm[i:i + 1], out["rdm"])
        d = dets[dets["cpi"] == i].copy()
        d["cpi"] = 0
        _compare_flags(d, out, 1, C, P, R, lib)
    # replicas of one distinct CPI agree bit for bit wherever they sit in the batch (chunk / slot boundaries included)
    for src in range(D):
        idx = np.flatnonzero(order == src)
        for j in idx[1:]:
            assert np.array_equal(rdm[idx[0]], rdm[j])


def test_chain_without_rdm_output_matches_across_chunks(lib):
    """rdm_out = NULL with several chunks in flight on the slot streams: each chunk needs its own RDM scratch (the range
    stage and the detection amplitudes read it).  Detections must equal those of the run that returns the RDM."""
    P, R, C, B = 64, 1024, 16, 6
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=6, r_lo=50, r_hi=R - 100)
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)

    def key(d):
        return np.sort(d, order=["cpi", "lane", "v", "r", "kind"])

    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=1, max_det=1 << 20) as ctx:
        rdm, dets_a, na = ctx.chain(raw, B, want_rdm=True)
        for _ in range(5):
            _, dets_b, nb = ctx.chain(raw, B, want_rdm=False)
            assert na == nb
            a, b = key(dets_a), key(dets_b)
            assert all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


@pytest.mark.parametrize("R,B,chunk", [(4096, 2, 0), (4096, 5, 2), (1000, 3, 0), (752, 1, 0), (188, 2, 0)])
def test_chain_single_pass_kernel(lib, monkeypatch, R, B, chunk):
    """RB200_ONEPASS=1: the single-pass kernel (unpack + PC + Doppler + 0-v + velocity CFAR with the pulse-compressed
    intermediate in shared memory, onepass_kernel.cu) against the oracle and against the default pipeline: ragged last
    tiles (R not a multiple of the 188-cell tile), a single tile, several chunks per call."""
    P, C = 64, 16
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=4, r_lo=20, r_hi=max(R - 80, 40))
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("single", ref), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=chunk, max_det=1 << 20) as ctx:
        rdm_d, dets_d, _ = ctx.chain(raw, B)
        launches_default = ctx.last_launch_count()
    monkeypatch.setenv("RB200_ONEPASS", "1")
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=chunk, max_det=1 << 20) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
        assert ctx.last_launch_count() <= launches_default         # de-interleave + single-pass kernel instead of PC + MTD (plus the range stage)
        rdm2, dets2, n2 = ctx.chain(raw, B)                        # the lane planes are reused between calls
        _, dets3, n3 = ctx.chain(raw, B, want_rdm=False)
    print("single-pass RDM rel err %.2e" % _close(rdm, out["rdm"]))
    _compare_flags(dets, out, B, C, P, R, lib)
    _close(rdm, rdm_d, tol=1e-5)
    assert np.array_equal(rdm, rdm2) and n == n2 == n3

    def key(d):
        return np.sort(d, order=["cpi", "lane", "v", "r", "kind"])

    a, b, c3 = key(dets), key(dets2), key(dets3)
    assert all(np.array_equal(a[f], b[f]) and np.array_equal(a[f], c3[f]) for f in a.dtype.names)


@pytest.mark.parametrize("R,B", [(4096, 2), (1000, 1)])
def test_chain_tensor_core_doppler_experiment(lib, monkeypatch, R, B):
    """RB200_MTD_TC=1 (+ RB200_NO_FUSED=1): the 64-point slow-time transform as a tcgen05 GEMM with a two-term bf16 split
    (mtd64_tc_kernel.cu, an experiment measured against the butterfly kernels): the RDM stays inside the 1e-4 gate
    (MP/fun_Process_MTD.m:20-26, MP/fun_0v_pressing.m:4-6) and the CFAR flags agree outside the near-threshold set."""
    P, C = 64, 16
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=4, r_lo=20, r_hi=max(R - 80, 40))
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("single", ref), cfar, near_tol=RTOL)
    monkeypatch.setenv("RB200_NO_FUSED", "1")
    monkeypatch.setenv("RB200_MTD_TC", "1")
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, max_det=1 << 20) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
    err = _close(rdm, out["rdm"])
    print("tensor-core Doppler RDM rel err %.2e" % err)
    assert err > 1e-7          # the bf16 split is visible: this really is the GEMM path, not the fp32 butterflies
    _compare_flags(dets, out, B, C, P, R, lib)


@pytest.mark.parametrize("env", [{"RB200_COEXIST": "1"}, {"RB200_NO_PCW": "1"}])
def test_chain_k1_variants_agree(lib, monkeypatch, env):
    """The K1 variants behind the experiment switches -- the 12-warp pcw_shared_kernel of RB200_COEXIST (with one Doppler CTA
    per SM) and the round-1 CTA-wide pc_fft_tma_kernel (RB200_NO_PCW) -- against the default warp-private pcw_kernel: the same
    overlap-save arithmetic (MP/fun_pulse_compression.m:16-22), so the RDM agrees to rounding and the detections are identical
    outside the near-threshold set of the oracle."""
    P, R, C, B = 64, 1500, 16, 3
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=4, r_lo=20, r_hi=R - 80)
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("single", ref), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=1, max_det=1 << 20) as ctx:
        rdm_d, dets_d, _ = ctx.chain(raw, B)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=1, max_det=1 << 20) as ctx:
        rdm, dets, _ = ctx.chain(raw, B)
    _close(rdm, out["rdm"])
    _close(rdm, rdm_d, tol=1e-5)
    _compare_flags(dets, out, B, C, P, R, lib)


@pytest.mark.parametrize("R,mti,use_cfar_default", [(2000, 30, True), (1000, 0, True), (1512, 30, False)])
def test_chain_p256_doppler_kernels_agree(lib, monkeypatch, R, mti, use_cfar_default):
    """P = 256: the persistent TMA-fed Doppler kernel (mtd256_tma_kernel: two 512-thread halves per SM that take turns on one
    staging tile) against the one-tile-per-CTA mtd_fast_kernel behind RB200_NO_TMA_MTD=1.  Both run mtd_fast_body on the same
    samples (MP/fun_Process_MTI.m:20-22, MP/fun_Process_MTD.m:20-30, CW/executeCFAR.m:28), so RDM, detection list and count
    must be bit-identical -- over several items per half (R = 2000: 63 tiles x 15 slabs), a ragged last tile (cols % 32 != 0),
    MTI on and off, and the run-time (ref, guard) CFAR variant."""
    P, C, B = 256, 5, 3
    ref = mcode.load_ref("refDBFDataMF1")
    raw = np.stack([synth.s3_cpi(i, P=P, R=R, C=C, ref=ref, seed0=7100, r_lo=100, r_hi=R - 200, n_targets=3)[0] for i in range(B)])
    cfg = dict(synth.S5_CFAR)
    if not use_cfar_default:
        cfg.update(refV=4, saveV=6)
    cfar = synth.cfar_tuple(cfg)

    def run():
        with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, mti_lag=mti, chunk_cpi=3, max_det=1 << 20) as ctx:
            rdm, dets, n = ctx.chain(raw, B)
        return rdm, np.sort(dets, order=["cpi", "lane", "v", "r", "kind"]), n

    rdm0, dets0, n0 = run()
    monkeypatch.setenv("RB200_NO_TMA_MTD", "1")
    rdm1, dets1, n1 = run()
    assert n0 == n1 and n0 > 0
    assert np.array_equal(rdm0, rdm1)
    assert all(np.array_equal(dets0[f], dets1[f]) for f in dets0.dtype.names)


@pytest.mark.parametrize("n1", [100, 74])
def test_chain_split_schedule_device_resident(lib, monkeypatch, n1):
    """RB200_SPLIT=n1: K1 (pcw_kernel) on n1 SMs and K2 (mtd64_tma_kernel) on the remaining SMs at the same time, K2 fetching a
    tile only when K1's per-CPI progress counter says the CPI is complete.  Same kernels, so RDM and detections of the
    device-resident call are bit-identical to the ordinary back-to-back schedule, over several chunks and repeated calls."""
    import torch
    P, R, C, B = 64, 1500, 16, 7
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=4, r_lo=20, r_hi=R - 80)
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    dev = torch.device("cuda", 0)
    raw_d = torch.from_numpy(np.ascontiguousarray(raw)).to(dev)

    def run(ctx):
        rdm_d = torch.zeros((B, C, P, R), dtype=torch.float32, device=dev)
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            ctx.chain_enqueue(raw_d.data_ptr(), B, rdm_d.data_ptr(), stream.cuda_stream)
            dets, n = ctx.chain_fetch()
        stream.synchronize()
        return rdm_d.cpu().numpy(), np.sort(dets, order=["cpi", "lane", "v", "r", "kind"]), n

    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=2, max_det=1 << 20) as ctx:
        rdm0, dets0, n0 = run(ctx)
    monkeypatch.setenv("RB200_SPLIT", str(n1))
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=2, max_det=1 << 20) as ctx:
        for _ in range(3):
            rdm1, dets1, n1_ = run(ctx)
            assert n1_ == n0 and np.array_equal(rdm1, rdm0)
            assert all(np.array_equal(dets1[f], dets0[f]) for f in dets0.dtype.names)


@pytest.mark.parametrize("R,use_stc,P", [(1032, False, 64), (1450, True, 64), (868, True, 32)])
def test_chain_pcw_kernel_multi_segment_plans(lib, monkeypatch, R, use_stc, P):
    """The warp-private K1 (pcw_kernel) on a THREE-segment waveform plan with 16 wire lanes and an even PRT length: FIR
    segment with its delay and rotation, two matched-filter segments with different spectra (h_off), tiles that touch the
    neighbouring segment (zeroed outside [0, in_len)), tiles before the PRT start (negative TMA coordinates) and the iSTC
    gains (MP/fun_lss_pulse_compression.m:24-37, MP/fun_iSTC.m:14) -- against the oracle.  RB200_PC_NT=256 keeps the 160-tap
    segment on 256-sample tiles (the plan would otherwise give it 4096-point tiles and K1 would fall back to pc_fft_kernel)."""
    monkeypatch.setenv("RB200_PC_NT", "256")
    C, B = 16, 2
    ref = mcode.load_ref("refDDCDataMF1")
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, ref=ref, n_targets=3, r_lo=20, r_hi=R - 80, seed0=4242)
    p2, p3 = mcode.load_pulse_literals()
    cfar = synth.cfar_tuple(synth.S3_CFAR) if P == 64 else (5, 7, 5.0, 0, 3, 2, 5.0, 0, 0, 1)
    stc = synth.s5_stc_curve()[: min(1025, R)] if use_stc else None
    out = vec.chain(raw, B, P, R, C, ("lss_mp", p2, p3), cfar, stc=stc, near_tol=RTOL)
    with lib.Context(0, n_prt=P, n_range=R, n_lanes=C, max_cpi=B, max_det=1 << 21) as ctx:
        ctx.set_waveform(lib.waveforms.segments_mp(R, p2, p3))
        ctx.set_cfar(*cfar)
        if stc is not None:
            ctx.set_stc(stc)
        rdm, dets, n = ctx.chain(raw, B)
    print("pcw multi-segment RDM rel err %.2e" % _close(rdm, out["rdm"]))
    _compare_flags(dets, out, B, C, P, R, lib)


def test_chain_single_pass_kernel_falls_back_outside_its_envelope(lib, monkeypatch):
    """With RB200_ONEPASS=1, configurations the single-pass kernel does not cover (13 lanes, a three-segment waveform, iSTC,
    R not a multiple of 4) silently take the slot pipeline and still match the oracle."""
    monkeypatch.setenv("RB200_ONEPASS", "1")
    P, R, C, B = 64, 1031, 13, 1
    rng = np.random.default_rng(3)
    raw = rng.integers(-300, 300, size=(B, P, R, C, 2), dtype=np.int16)
    p2, p3 = mcode.load_pulse_literals()
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("lss_mp", p2, p3), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_mp(R, p2, p3), cfar) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
    _close(rdm, out["rdm"])
    _compare_flags(dets, out, B, C, P, R, lib)


def test_chain_detection_overflow_is_reported(lib):
    P, R, C, B = 64, 512, 2, 1
    raw = np.zeros((B, P, R, C, 2), dtype=np.int16)             # all-zero RDM: 0 >= 0 flags every cell
    ref = mcode.load_ref("refDDCDataMF1")
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), synth.cfar_tuple(synth.S3_CFAR), max_det=1000) as ctx:
        with pytest.raises(lib.DetectionOverflow):
            ctx.chain(raw, B)
        rdm, dets, n = ctx.chain(raw, B, allow_overflow=True)
        assert n > 1000 and len(dets) == 1000


@pytest.mark.parametrize("case", [
    # P,  R,    C, B, cfar (refR,gR,T_R,mR, refV,gV,T_V,mV, n0, rflag), mti, zero_div, chunk
    dict(P=64, R=1000, C=1, B=3, cfar=(5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1), mti=0, zdiv=150, chunk=2),    # single lane, ragged R, fused path
    dict(P=64, R=777, C=20, B=1, cfar=(5, 7, 4.0, 1, 5, 7, 4.0, 1, 0, 1), mti=0, zdiv=20, chunk=1),     # >16 lanes (two lane groups), SO-CFAR, odd R
    dict(P=64, R=512, C=3, B=2, cfar=(3, 2, 4.0, 0, 4, 3, 4.0, 0, 2, 1), mti=0, zdiv=150, chunk=1),     # non-default windows + n0 -> unfused path
    dict(P=64, R=640, C=2, B=2, cfar=(5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 0), mti=0, zdiv=150, chunk=2),     # range stage off
    dict(P=64, R=512, C=2, B=1, cfar=(5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1), mti=30, zdiv=150, chunk=1),    # MTI on P=64 -> shared-memory MTD path
    dict(P=128, R=384, C=2, B=1, cfar=(5, 7, 5.0, 0, 5, 7, 5.0, 0, 1, 1), mti=0, zdiv=150, chunk=1),    # generic Stockham Doppler length
    dict(P=48, R=300, C=2, B=2, cfar=(5, 7, 5.0, 0, 4, 3, 5.0, 0, 0, 1), mti=0, zdiv=0, chunk=2),       # P = 2^4*3, no zero-velocity mask
    dict(P=256, R=333, C=2, B=2, cfar=(4, 2, 5.0, 1, 6, 3, 5.0, 1, 3, 1), mti=0, zdiv=150, chunk=1),    # P=256 fused V-CFAR, run-time windows, SO, ragged R
    dict(P=256, R=288, C=1, B=1, cfar=(5, 7, 6.0, 0, 5, 7, 6.0, 0, 0, 0), mti=0, zdiv=20, chunk=1),     # P=256 fused V-CFAR, default windows, range stage off
    dict(P=256, R=224, C=2, B=1, cfar=(5, 7, 6.0, 0, 9, 20, 6.0, 0, 40, 1), mti=30, zdiv=150, chunk=1),  # wide guard band + large n0: edge substitution on most rows
])
def test_chain_variants(lib, case):
    P, R, C, B = case["P"], case["R"], case["C"], case["B"]
    ref = mcode.load_ref("refDDCDataMF1")
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=3, r_lo=20, r_hi=R - 80)
    if case["zdiv"]:
        out = vec.chain(raw, B, P, R, C, ("single", ref), case["cfar"], zero_div=case["zdiv"], mti_lag=case["mti"], near_tol=RTOL)
    else:
        x = vec.single_pc(vec.unpack_wire(raw, B, P, R, C), ref)
        rdm0 = vec.process_mtd(x)
        res = vec.execute_cfar(rdm0, *case["cfar"], near_tol=RTOL)
        out = {"rdm": rdm0, "flag": res[0], "flagV": res[1], "near": res[2], "nearV": res[3]}
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), case["cfar"], mti_lag=case["mti"], zero_v_div=case["zdiv"],
                    chunk_cpi=case["chunk"], max_det=1 << 20) as ctx:
        rdm, dets, n = ctx.chain(raw, B)
    _close(rdm, out["rdm"])
    _compare_flags(dets, out, B, C, P, R, lib)
    if not case["cfar"][-1]:        # rCFARDetect_Flag = 0: every velocity hit is also the final flag (executeCFAR.m:91)
        assert (dets["kind"] == (lib.DET_V | lib.DET_2D)).all()


def test_chain_argument_errors(lib):
    ref = mcode.load_ref("refDDCDataMF1")
    with lib.Context(0, n_prt=64, n_range=256, n_lanes=2, max_cpi=2) as ctx:
        raw = np.zeros((1, 64, 256, 2, 2), dtype=np.int16)
        with pytest.raises(lib.RadarB200Error) as e:          # no waveform yet
            ctx.chain(raw, 1)
        assert e.value.status == 5
        ctx.set_waveform(lib.waveforms.segments_single(512, ref))   # plan longer than the PRT
        with pytest.raises(lib.MatlabIndexError):
            ctx.chain(raw, 1)
        ctx.set_waveform(lib.waveforms.segments_single(256, ref))
        with pytest.raises(lib.RadarB200Error):               # n_cpi > max_cpi
            ctx.chain(np.zeros((3, 64, 256, 2, 2), dtype=np.int16), 3)
    with lib.Context(0, n_prt=16, n_range=256, n_lanes=1, max_cpi=1) as ctx:   # 15 tested rows < 2*(5+7)
        ctx.set_waveform(lib.waveforms.segments_single(256, ref))
        with pytest.raises(lib.MatlabIndexError):
            ctx.chain(np.ones((1, 16, 256, 1, 2), dtype=np.int16), 1)


def test_chain_device_resident_fused_persistent_kernel(lib):
    """The opt-in fused persistent kernel (RB200_MEGA=1: PC and MTD roles, L2 ring of 3 CPIs) on device-resident
    buffers.  Five CPIs wrap the ring; results must agree with the chunked path and match the oracle."""
    import os
    import torch
    os.environ["RB200_MEGA"] = "1"
    P, R, C, B = 64, 4096, 16, 5
    raw, _ = synth.s3_batch(B)
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    dev = torch.device("cuda", 0)
    raw_d = torch.from_numpy(raw).to(dev)
    rdm_d = torch.zeros((B, C, P, R), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, max_det=1 << 18) as ctx:
        try:
            with torch.cuda.stream(stream):
                ctx.chain_enqueue(raw_d.data_ptr(), B, rdm_d.data_ptr(), stream.cuda_stream)
                dets_m, n_m = ctx.chain_fetch()
                stream.synchronize()
        finally:
            del os.environ["RB200_MEGA"]
        assert ctx.last_launch_count() == 2           # chain64_kernel + cfar_r64_kernel
        rdm_m = rdm_d.cpu().numpy()
        rdm_h, dets_h, n_h = ctx.chain(raw, B)        # host buffers -> chunked slot pipeline
    rel = np.max(np.abs(rdm_m - rdm_h)) / np.max(rdm_h)
    print("fused-persistent vs chunked RDM: max rel diff %.3e, differing cells %d" % (rel, int((rdm_m != rdm_h).sum())))
    assert rel < 1e-6      # same algorithm; instruction scheduling / FMA contraction may differ between the two kernels

    def key(d):
        return np.sort(d, order=["cpi", "lane", "v", "r", "kind"])

    a, b = key(dets_m), key(dets_h)
    assert n_m == n_h and len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in ("cpi", "lane", "v", "r", "kind"))
    out = vec.chain(raw[:2], 2, P, R, C, ("single", ref), cfar, near_tol=RTOL)
    _close(rdm_m[:2], out["rdm"])
    _compare_flags(dets_m[dets_m["cpi"] < 2], out, 2, C, P, R, lib)


def test_chain_is_deterministic_across_runs(lib):
    """Repeated runs of the same device-resident batch (two chunks on two slot streams, TMA-staged kernels running
    concurrently) give bit-identical RDMs and identical detection sets.  Regression test for a cross-proxy WAR race
    (generic-proxy tile reads vs the TMA refill) that showed up as 32-column blocks of wrong RDM values."""
    import torch
    P, R, C, B = 64, 4096, 16, 16
    raw = np.concatenate([synth.s3_batch(4)[0]] * 4, axis=0)
    ref = mcode.load_ref("refDDCDataMF1")
    dev = torch.device("cuda", 0)
    raw_d = torch.from_numpy(raw).to(dev)
    rdm_d = torch.zeros((B, C, P, R), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)

    def key(d):
        return np.sort(d, order=["cpi", "lane", "v", "r", "kind"])

    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), synth.cfar_tuple(synth.S3_CFAR), max_det=1 << 19,
                    chunk_cpi=8) as ctx:
        first = None
        for it in range(25):
            with torch.cuda.stream(stream):
                rdm_d.zero_()
                ctx.chain_enqueue(raw_d.data_ptr(), B, rdm_d.data_ptr(), stream.cuda_stream)
                dets, n = ctx.chain_fetch()          # ordered after the enqueue on the same stream
                stream.synchronize()
            cur = (rdm_d.cpu().numpy(), key(dets))
            if first is None:
                first = cur
                assert n > 0
                continue
            assert np.array_equal(cur[0], first[0]), "run %d: %d RDM cells differ" % (it, int((cur[0] != first[0]).sum()))
            assert len(cur[1]) == len(first[1]) and all(np.array_equal(cur[1][f], first[1][f]) for f in cur[1].dtype.names)


def test_chain_with_dbf_weighting(lib):
    """16 DDC channels -> 13 beams (sig_C * W.') fused with the unpack, then the usual chain per beam."""
    P, R, C, NB, B = 64, 1024, 16, 13, 2
    rng = np.random.default_rng(31)
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=3, r_lo=50, r_hi=R - 100)
    W = (rng.normal(size=(NB, C)) + 1j * rng.normal(size=(NB, C))) / 4.0
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(raw, B, P, R, C, ("single", ref), cfar, near_tol=RTOL, dbf=W)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=1) as ctx:
        ctx.set_dbf(W)
        rdm, dets, n = ctx.chain(raw, B)
        assert rdm.shape == (B, NB, P, R)
        pc = ctx.debug_fetch_pc(0)
        _close(pc, out["pc"][B - 1])
        _close(rdm, out["rdm"])
        _compare_flags(dets, out, B, NB, P, R, lib)
        ctx.set_dbf(None)                                    # back to per-channel processing
        rdm16, _, _ = ctx.chain(raw, B)
        assert rdm16.shape == (B, C, P, R)
        _close(rdm16, vec.chain(raw, B, P, R, C, ("single", ref), cfar)["rdm"])


@pytest.mark.parametrize("ch,n,n_prt", [(13, 37, 3), (4, 16, 2), (16, 5, 1)])
def test_unpack_dbf24_bit_exact(lib, ch, n, n_prt):
    rng = np.random.default_rng(ch)
    sig, pad, osp = mcode.dbf24_payload_size(n, ch)
    payloads = rng.integers(0, 256, size=(n_prt, sig + pad), dtype=np.uint8)
    payloads[0, 0:6] = (0, 0, 0x80, 1, 0, 0x80)                  # 0x800000 -> +2^23 (quirk), 0x800001 -> -(2^23-1)
    with lib.Context(0) as ctx:
        got = ctx.unpack_dbf24(payloads, n_prt, n, ch)
    want = np.stack([mcode.unpack_dbf24(payloads[p], n, ch) for p in range(n_prt)], axis=0)      # prt, sample, col
    assert got.shape == (want.shape[2], n_prt, n)
    assert np.array_equal(got.astype(np.complex128), want.transpose(2, 0, 1))
    assert got[0, 0, 0] == 8388608 - 8388607j


def test_sliding_window_reuse_matches_per_window_calls(lib):
    """the 4-window loop of MP/main_produce_dataset_win_xzr.m:24-38 with the pulse compression
    done once over the concatenated frames equals fun_MTD_produce on every window slice."""
    rng = np.random.default_rng(41)
    win_len, win_size, n = 256, 4, 1031
    echo_win = np.rint(150 * _rand_c(rng, 2 * win_len, n))                     # [frame N; frame N+1]
    got = lib.fun_MTD_produce_windows(echo_win, win_len, win_size)
    assert got.shape == (win_size, win_len, n)
    for i in range(win_size):
        s0 = mcode.mround(i * win_len / win_size)                              # round(i*1536/win_size)+1 (1-based)
        want = mcode.fun_MTD_produce_mp(echo_win[s0:s0 + win_len]) if i == 1 else None
        one = lib.fun_MTD_produce(echo_win[s0:s0 + win_len])
        assert np.array_equal(got[i], one)                                     # same kernels, same data -> same bits
        if want is not None:
            _close(got[i], want)
    with lib.Context(0) as ctx:
        ctx.set_waveform(lib.waveforms.segments_mp(n, lib.waveforms.PULSE2, lib.waveforms.PULSE3))
        with pytest.raises(lib.MatlabIndexError):
            ctx.mtd_produce_windows(echo_win, win_len, [0, 2 * win_len - 10])


def test_motionParaMeasure_matches_oracle(lib):
    """post-CFAR measurement on the flagged cells (spline refinement + monopulse elevation)."""
    rng = np.random.default_rng(51)
    V, R, n0 = 64, 200, 3
    s = rng.rayleigh(1.0, size=(V, R)) + 0.5
    d = rng.normal(size=(V, R))
    flags = np.zeros((V, R))
    for _ in range(25):
        v, r = int(rng.integers(n0 + 1, V - n0)), int(rng.integers(0, R))
        s[max(v - 2, 0):v + 3, max(r - 2, 0):r + 3] += 30 * rng.random()
        s[v, r] += 40
        flags[v, r] = 1
    flags[n0 + 1, 0] = flags[V - n0 - 1, R - 1] = flags[n0 + 2, 1] = 1            # edge windows on both axes
    rScale = 6.0 * np.arange(R) + 100.0
    vScale = 0.27 * (V // 2 - np.arange(V))
    kValues = 10 + rng.random((11, 12))
    for extra, rt, vt in ((2, 8, 4), (1, 5, 3), (4, 2, 2)):
        args = (extra, rScale, 6.0, rt, vScale, 0.27, vt, kValues, 7, 3.0, 4, 0.3, -0.1, n0)
        got = lib.motionParaMeasure(s, d, flags, *args)
        want = mcode.motionParaMeasure(s, d, flags, *args)
        assert len(got) == 3 and got[0].shape == want[0].shape == (int(flags.sum()),)
        for g, w in zip(got, want):
            assert np.allclose(g, w, rtol=1e-11, atol=1e-9)
    empty = lib.motionParaMeasure(s, d, np.zeros((V, R)), *args)
    assert all(e.shape == (0,) for e in empty)
    with pytest.raises(lib.MatlabIndexError):
        lib.motionParaMeasure(s, d, flags, 2, rScale, 6.0, 8, vScale, 0.27, 4, kValues, 12, 3.0, 4, 0, 0, n0)       # beamPosNum+1 > 12


def test_capture_files_to_detections_end_to_end(lib, tmp_path):
    """f2 + the hot path: framed capture files -> C++ reader -> (host buffer) -> rb200_chain_i16 -> detections, against
    the oracle's frame parser + chain on the same bytes."""
    from radar_signal_process_b200.reader import FrameReader
    P, R, C, B = 64, 512, 4, 2
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=2, r_lo=40, r_hi=R - 100)
    blob = b"".join(synth.frame_prt(raw[b, p], frame_no=b, prt_no=p, channel_num=C) for b in range(B) for p in range(P))
    cut = len(blob) // 3 + 17
    (tmp_path / "1.000001.bin").write_bytes(blob[:cut])
    (tmp_path / "1.000002.bin").write_bytes(blob[cut:])
    rd = FrameReader(tmp_path)
    batch = np.zeros((B, P, R, C, 2), dtype=np.int16)
    for b in range(B):
        _, meta, nread, eos = rd.next_frame(P, R, C, out=batch[b])
        assert nread == P and not eos and np.all(meta["frame_no"] == b)
    assert np.array_equal(batch, raw)
    ref = mcode.load_ref("refDDCDataMF1")
    cfar = synth.cfar_tuple(synth.S3_CFAR)
    out = vec.chain(batch, B, P, R, C, ("single", ref), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_single(R, ref), cfar) as ctx:
        rdm, dets, n = ctx.chain(batch, B)
    _close(rdm, out["rdm"])
    _compare_flags(dets, out, B, C, P, R, lib)


def _dmx_inputs(P, n_range, n_short, seed):
    rng = np.random.default_rng(seed)
    ref = mcode.load_ref("refDDCDataMF1")
    mf = mcode.dmx_match_filter(ref)
    beams = []
    for b in range(2):
        x = np.round(rng.normal(0, 40, (P, n_range))) + 1j * np.round(rng.normal(0, 40, (P, n_range)))
        # a long-pulse echo near the end of the PRT (wraps around in the circular compression) and one in the middle
        for r0, dop, amp in ((n_range - 30, 0.11, 300.0), (n_short + 100, -0.23, 200.0 * (1 + b))):
            n = min(ref.size, n_range - r0)
            x[:, r0:r0 + n] += amp * ref[None, :n] * np.exp(2j * np.pi * dop * np.arange(P))[:, None]
        if n_short:
            x[:, 20] += 500.0 * np.exp(2j * np.pi * 0.05 * np.arange(P))      # short-pulse return
        beams.append(x)
    return beams[0], beams[1], mf


@pytest.mark.parametrize("P,n_range,n_short,fft_num,mtd_fft,n0", [
    (96, 566, 62, 512, 128, 3),        # reduced slow-time length, script geometry in range (62 + 504 -> 512)
    (1536, 566, 62, 512, 2048, 27),    # the script's own sizes (prtNum 1536, FFT_num 512, mtd_FFT_num 2048)
    (40, 200, 0, 256, 40, 0),          # no short pulse, no Doppler zero padding
])
def test_dmx_frame_matches_oracle(lib, P, n_range, n_short, fft_num, mtd_fft, n0):
    """f4: DMX script variant -- FIR short pulse, circular matched filter, Hamming zero-padded MTD, sum / difference."""
    left, right, mf = _dmx_inputs(P, n_range, n_short, seed=P + n_range)
    win = mcode.hamming(P)
    want = mcode.dmx_frame(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0)
    with lib.Context(0, n_prt=8, n_range=64, n_lanes=1) as ctx:
        got = ctx.dmx_process(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0)
        again = ctx.dmx_process(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0)   # cached plan
    for g, a, w in zip(got, again, want):
        if n_short == 0 and w.shape[1] == 0:
            assert g is None
            continue
        assert g.shape == w.shape
        np.testing.assert_array_equal(g, a)
        scale = np.abs(w).max()
        assert np.abs(g - w).max() <= RTOL * scale, (np.abs(g - w).max(), scale)
    # blanked rows are exactly zero in the sums and untouched in the differences
    assert np.all(got[2][:n0 + 1] == 0) and np.all(got[2][mtd_fft - n0:] == 0)
    assert np.any(got[3][:n0 + 1] != 0)


def test_dmx_sums_feed_execute_cfar_like_the_script(lib):
    """DMX script :468-472: executeCFAR on the blanked long-pulse sum with n0 = MTD_0_num."""
    P, n_range, n_short, fft_num, mtd_fft, n0 = 96, 566, 62, 512, 128, 3
    left, right, mf = _dmx_inputs(P, n_range, n_short, seed=77)
    win = mcode.hamming(P)
    want = mcode.dmx_frame(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0)
    with lib.Context(0, n_prt=8, n_range=64, n_lanes=1) as ctx:
        got = ctx.dmx_process(left, right, n_short, mcode.FILTER_COEF_INT, mf, fft_num, win, mtd_fft, n0)
    f, fv = lib.executeCFAR(got[2], 5, 7, 7.0, 0, 5, 7, 7.0, 0, n0, 1)
    wf, wfv = mcode.executeCFAR(got[2], 5, 7, 7.0, 0, 5, 7, 7.0, 0, n0, 1)
    np.testing.assert_array_equal(f, wf)
    np.testing.assert_array_equal(fv, wfv)
    assert fv.sum() > 0


def test_dmx_argument_errors(lib):
    left, right, mf = _dmx_inputs(16, 120, 10, seed=1)
    win = mcode.hamming(16)
    with lib.Context(0, n_prt=8, n_range=64, n_lanes=1) as ctx:
        with pytest.raises(lib.RadarB200Error):       # FFT_num not a supported size
            ctx.dmx_process(left, right, 10, mcode.FILTER_COEF_INT, mf, 300, win, 16, 0)
        with pytest.raises(lib.RadarB200Error):       # long pulse longer than FFT_num (fft(x, n) would truncate)
            ctx.dmx_process(left, right, 10, mcode.FILTER_COEF_INT, mf[:20], 256, np.ones(16), 8, 0)
        with pytest.raises(lib.RadarB200Error):       # blanking wider than the Doppler axis
            ctx.dmx_process(left, right, 10, mcode.FILTER_COEF_INT, mf, 256, win, 16, 9)


@pytest.mark.parametrize("P,R,n_ch,B,chunk", [(64, 512, 13, 3, 2), (256, 300, 4, 1, 1)])
def test_chain_on_dbf24_payloads(lib, P, R, n_ch, B, chunk):
    """f1 (batched): DBF-type 24-bit frames straight into the chain; lanes = the complex columns the reference keeps."""
    ref = mcode.load_ref("refDBFDataMF1")
    rng = np.random.default_rng(P + n_ch)
    sig, pad, osp = mcode.dbf24_payload_size(R, n_ch)
    ncol = ((n_ch * 6 + osp) // 3) // 2
    lanes = np.round(rng.normal(0, 3000, (B, P, R, ncol))) + 1j * np.round(rng.normal(0, 3000, (B, P, R, ncol)))
    for b in range(B):
        for col in range(ncol):
            r0, dop = 40 + 17 * col + 5 * b, 0.07 * (col + 1) - 0.3
            echo = (30000.0 / np.abs(ref).max()) * ref[None, :] * np.exp(2j * np.pi * dop * np.arange(P))[:, None]
            lanes[b, :, r0:r0 + ref.size, col] += np.round(echo.real) + 1j * np.round(echo.imag)
    lanes[0, 0, 0, 0] = -8388607 + 8388607j                                   # extreme codes survive the round trip
    payload = np.stack([synth.to_dbf24(lanes[b], n_ch) for b in range(B)])       # [cpi][prt][bytes]
    decoded = np.stack([np.stack([mcode.unpack_dbf24(payload[b, p], R, n_ch) for p in range(P)]) for b in range(B)])
    assert np.array_equal(decoded, lanes)
    cfar = (5, 7, 6.0, 0, 5, 7, 6.0, 0, 0, 1)
    out = vec.chain_lanes(lanes.transpose(0, 3, 1, 2), ("single", ref), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, ncol, B, lib.waveforms.segments_single(R, ref), cfar, chunk_cpi=chunk, max_det=1 << 20) as ctx:
        rdm, dets, n = ctx.chain_dbf24(payload, n_ch, B)
        with pytest.raises(lib.RadarB200Error):                                 # column count must match n_lanes
            ctx.chain_dbf24(payload, n_ch + 3, B)
    _close(rdm, out["rdm"])
    _compare_flags(dets, out, B, ncol, P, R, lib)


def test_chain_S5_full_size_all_lanes_against_oracle(lib):
    """Configuration S5 at BASELINE's full size (256 PRT x 16384 range x 16 lanes, refDBFDataMF1, iSTC + MTI): the oracle
    is run lane by lane on all sixteen lanes (all range cells) and compared cell by cell."""
    P, R, C = 256, 16384, 16
    ref = mcode.load_ref("refDBFDataMF1")
    raw, _ = synth.s3_cpi(0, P=P, R=R, C=C, ref=ref, seed0=5000, r_lo=100, r_hi=R - 200, exclude=(-3, -2, -1, 0, 1, 2, 3))
    raw = raw[None]
    stc = synth.s5_stc_curve()
    cfar = synth.cfar_tuple(synth.S5_CFAR)
    with _chain_ctx(lib, P, R, C, 1, lib.waveforms.segments_single(R, ref), cfar, mti_lag=30, max_det=1 << 20) as ctx:
        ctx.set_stc(stc)
        rdm, dets, n = ctx.chain(raw, 1)
    assert np.all(rdm[:, :, 125:130, :] == 0) and n > 0
    for lane in range(C):
        sub = np.ascontiguousarray(raw[:, :, :, lane:lane + 1, :])
        out = vec.chain(sub, 1, P, R, 1, ("single", ref), cfar, stc=stc, mti_lag=30, near_tol=RTOL)
        _close(rdm[:, lane:lane + 1], out["rdm"])
        d = dets[dets["lane"] == lane].copy()
        d["lane"] = 0
        _compare_flags(d, out, 1, 1, P, R, lib)


def test_dbf24_capture_files_to_detections_end_to_end(lib, tmp_path):
    """f2 + f1 + the hot path on DBF-type captures: framed files (data_type 2) -> C++ reader -> rb200_chain_dbf24 -> detections,
    against the oracle's 24-bit decoder + chain on the same bytes."""
    from radar_signal_process_b200.reader import FrameReader
    P, R, n_ch, B = 64, 300, 13, 2
    ref = mcode.load_ref("refDBFDataMF1")
    rng = np.random.default_rng(21)
    sig, pad, osp = mcode.dbf24_payload_size(R, n_ch)
    ncol = ((n_ch * 6 + osp) // 3) // 2
    lanes = np.round(rng.normal(0, 2000, (B, P, R, ncol))) + 1j * np.round(rng.normal(0, 2000, (B, P, R, ncol)))
    for b in range(B):
        echo = (25000.0 / np.abs(ref).max()) * ref[None, :] * np.exp(2j * np.pi * (0.13 + 0.2 * b) * np.arange(P))[:, None]
        lanes[b, :, 90:90 + ref.size, 3 + b] += np.round(echo.real) + 1j * np.round(echo.imag)
    blob = b"".join(synth.frame_prt_dbf24(synth.to_dbf24(lanes[b, p:p + 1], n_ch)[0], R, frame_no=b, prt_no=p, channel_num=n_ch)
                    for b in range(B) for p in range(P))
    cut = len(blob) // 2 + 33
    (tmp_path / "1.000001.bin").write_bytes(blob[:cut])
    (tmp_path / "1.000002.bin").write_bytes(blob[cut:])
    rd = FrameReader(tmp_path)
    frames = []
    for b in range(B):
        payload, meta, nread, eos = rd.next_frame_dbf24(P, R, n_ch)
        assert nread == P and not eos and np.all(meta["frame_no"] == b)
        frames.append(payload.copy())
    batch = np.stack(frames)                                                     # [cpi][prt][padded bytes]
    cfar = (5, 7, 6.0, 0, 5, 7, 6.0, 0, 0, 1)
    out = vec.chain_lanes(lanes.transpose(0, 3, 1, 2), ("single", ref), cfar, near_tol=RTOL)
    with _chain_ctx(lib, P, R, ncol, B, lib.waveforms.segments_single(R, ref), cfar, max_det=1 << 20) as ctx:
        rdm, dets, n = ctx.chain_dbf24(batch, n_ch, B)
    _close(rdm, out["rdm"])
    _compare_flags(dets, out, B, ncol, P, R, lib)
    flag, _ = lib.dets_to_flags(dets, B, ncol, P, R)
    assert flag[0, 3].sum() > 0 and flag[1, 4].sum() > 0


def _segmented_cfar_oracle(rdm, cfar, segments):
    """fun_CFARflag (CW/main_cfar.m:142-161) with the dense oracle: executeCFAR per range segment, zeros elsewhere."""
    out = {k: np.zeros(rdm.shape, dtype=bool) for k in ("flag", "flagV", "near", "nearV")}
    for lo, hi in segments:
        f, fv, near, nearv = vec.execute_cfar(rdm[..., lo:hi], *cfar, near_tol=RTOL)
        out["flag"][..., lo:hi], out["flagV"][..., lo:hi] = f.astype(bool), fv.astype(bool)
        out["near"][..., lo:hi], out["nearV"][..., lo:hi] = near, nearv
    return out


@pytest.mark.parametrize("P,C,B,mti", [(64, 16, 2, 0), (64, 3, 1, 0), (256, 2, 1, 30), (96, 2, 1, 0)])
def test_chain_cfar_range_segments_like_fun_CFARflag(lib, P, C, B, mti):
    """The three-segment waveform with the CFAR confined to each segment (CW/main_cfar.m:56-58,142-161): range windows do not
    straddle segment borders, columns beyond the last segment stay empty.  Fused (P=64, 16 lanes), shared-memory and generic paths."""
    R = 1031
    segments = [(0, 82), (82, 318), (318, 868)]
    p2, p3 = mcode.load_pulse_literals()
    raw, _ = synth.s3_batch(B, P=P, R=R, C=C, n_targets=4, r_lo=20, r_hi=R - 80, seed0=77 + P)
    cfar = (5, 7, 5.0, 0, 5, 7, 5.0, 0, 0, 1)
    base = vec.chain(raw, B, P, R, C, ("lss_mp", p2, p3), cfar, mti_lag=mti, near_tol=RTOL)
    want = _segmented_cfar_oracle(base["rdm"], cfar, segments)
    with _chain_ctx(lib, P, R, C, B, lib.waveforms.segments_mp(R, p2, p3), cfar, mti_lag=mti, max_det=1 << 20) as ctx:
        ctx.set_cfar_segments(segments)
        rdm, dets, n = ctx.chain(raw, B)
        _close(rdm, base["rdm"])
        _compare_flags(dets, want, B, C, P, R, lib)
        assert not np.any(dets["r"] >= 868)
        # the un-segmented result differs (windows straddle the borders) and is restored by an empty list
        ctx.set_cfar_segments([])
        _, dets0, _ = ctx.chain(raw, B)
        _compare_flags(dets0, base, B, C, P, R, lib)
        assert np.any(dets0["r"] >= 868)
        # a segment shorter than the range windows: Function_CFAR1D_sub_fixCells indexes out of bounds -- but, as in the
        # M-code, only when a velocity hit falls into that segment (the range stage runs per hit, executeCFAR.m:45-61)
        ctx.set_cfar_segments([(0, 20), (20, 868)])
        # cells 8..11 of a 20-cell segment fit neither window (ref 5 + guard 7 on both sides): hits at columns 7..12 test one
        if want["flagV"][..., 7:13].any():
            with pytest.raises(lib.MatlabIndexError):
                ctx.chain(raw, B)
        else:
            ctx.chain(raw, B)
        with pytest.raises(lib.RadarB200Error):
            ctx.set_cfar_segments([(0, 100), (50, 200)])
