"""Seeded synthetic inputs S1..S5 and the int16 DDC wire writer (TEST ORACLE).

The reference's only recorded frame (``echoData_Frame.mat``) is absent from the tree
(``.MISSING_LARGE_BLOBS``) and ``fun_SimulateTarget`` is missing (MP/main.m:188), so every
configuration is driven by these deterministic generators.  Wire layout follows
FrameDataRead_xzr.m:109,150-156: per PRT ``n*ch*2`` little-endian int16 ordered
``[range][channel][I,Q]``; a CPI batch is ``[cpi][prt][range][channel][I,Q]`` with no framing.
"""
import numpy as np

from . import mcode

C_LIGHT = 2.99792458e8


def to_wire(x):
    """complex ``[cpi][lane][prt][range]`` (integer-valued or not) -> saturated int16 wire array
    ``[cpi][prt][range][lane][2]``."""
    x = np.asarray(x)
    a = np.stack([x.real, x.imag], axis=-1)              # cpi lane prt range 2
    a = np.rint(a).clip(-32768, 32767).astype(np.int16)
    return np.ascontiguousarray(a.transpose(0, 2, 3, 1, 4))


def frame_prt(payload_i16, frame_no=0, prt_no=0, channel_num=16, data_type=1, servo=0):
    """Wrap one PRT payload in the 64 B head / 128 B realtime / payload(+pad to 64 B) / 64 B tail
    framing parsed at FrameDataRead_xzr.m:62-119,184 (bytes)."""
    payload = np.ascontiguousarray(payload_i16, dtype="<i2").tobytes()
    n = len(payload) // (channel_num * 4)
    head = np.zeros(16, dtype="<u4")
    head[0] = frame_no
    head[2] = prt_no & 0xFFFF
    head[3] = channel_num & 0xFF
    head[4] = servo & 0xFFFF
    head[6] = n
    head[7] = data_type & 0xFF
    sig, pad = mcode.ddc_payload_size(n, channel_num)
    assert sig == len(payload)
    return head.tobytes() + bytes(128) + payload + bytes(pad) + bytes(64)


def _snr_amp(snr_db, ref, window, noise_sigma):
    """Echo scale so that the post-PC, post-Doppler peak SNR equals ``snr_db``."""
    e_ref = float(np.sum(np.abs(ref) ** 2))
    sw, sw2 = float(window.sum()), float((window ** 2).sum())
    snr = 10.0 ** (snr_db / 10.0)
    return np.sqrt(snr * 2.0 * noise_sigma ** 2 * sw2 / (e_ref * sw * sw))


def s3_cpi(cpi, P=64, R=4096, C=16, ref=None, n_targets=8, noise_sigma=64.0, seed0=1234,
           r_lo=100, r_hi=3900, exclude=(-1, 0)):
    """One S3 CPI in wire layout ``[prt][range][lane][2]`` int16 plus its target table."""
    if ref is None:
        ref = mcode.load_ref("refDDCDataMF1")
    rng = np.random.default_rng(seed0 + cpi)
    L = ref.shape[0]
    w = mcode.kaiser(P, 8.0)
    a = rng.normal(0.0, noise_sigma, size=(P, R, C, 2))
    np.rint(a, out=a)
    p = np.arange(P)
    targets = []
    bins = [k for k in range(-(P // 2), P - P // 2) if k not in exclude]
    for lane in range(C):
        for _ in range(n_targets):
            r0 = int(rng.integers(r_lo, r_hi + 1))
            k = int(bins[int(rng.integers(0, len(bins)))])
            snr_db = float(rng.uniform(15.0, 30.0))
            amp = _snr_amp(snr_db, ref, w, noise_sigma)
            ph = np.exp(2j * np.pi * k * p / P)
            n_valid = min(L, R - r0)
            echo = amp * ph[:, None] * ref[None, :n_valid]
            a[:, r0:r0 + n_valid, lane, 0] += echo.real
            a[:, r0:r0 + n_valid, lane, 1] += echo.imag
            targets.append((lane, r0, k, snr_db))
    np.rint(a, out=a)
    return a.clip(-32768, 32767).astype(np.int16), targets


def s3_batch(n_cpi, first_cpi=0, **kw):
    out = []
    tg = []
    for c in range(first_cpi, first_cpi + n_cpi):
        a, t = s3_cpi(c, **kw)
        out.append(a)
        tg.append(t)
    return np.stack(out, axis=0), tg


def s5_stc_curve():
    """Synthetic STC curve (dB): linear 30 -> 0 over 1025 entries (mirrors MP/fun_iSTC.m:6-9)."""
    r = np.arange(1025, dtype=np.float64)
    return 30.0 * (1.0 - r / 1024.0)


def s5_cpi(cpi, P=256, R=16384, C=16, ref=None, n_targets=8, noise_sigma=64.0):
    if ref is None:
        ref = mcode.load_ref("refDBFDataMF1")
    return s3_cpi(cpi, P=P, R=R, C=C, ref=ref, n_targets=n_targets, noise_sigma=noise_sigma,
                  seed0=5000, r_lo=100, r_hi=R - 200, exclude=(-3, -2, -1, 0, 1, 2, 3))


def s1_frame(beam=0, P=1536, R=1031, seed=20220420):
    """S1 stand-in for echoData_Frame_<beam>: complex double P x R, integer-valued."""
    rng = np.random.default_rng(seed + beam)
    x = np.rint(rng.normal(0.0, 200.0, size=(P, R))) + 1j * np.rint(rng.normal(0.0, 200.0, size=(P, R)))
    pulse1 = mcode.pulse1_mp()
    pulse2, pulse3 = mcode.load_pulse_literals()
    p = np.arange(P)
    A = 4000.0
    for pulse, col1, k in ((pulse1, 40, 37), (pulse2, 83 + 100, -112), (pulse3, 325 + 300, 300)):
        s = np.asarray(pulse, dtype=np.complex128) * (A / np.abs(pulse).max())
        ph = np.exp(2j * np.pi * k * p / P)
        x[:, col1 - 1:col1 - 1 + s.shape[0]] += ph[:, None] * s[None, :]
    return np.rint(x.real) + 1j * np.rint(x.imag)


def s2_frame(P=8, R=1024, seed=8, scr_db=10.0, velocity=-5.7, range_m=320.0):
    """S2: README worked example size, target per MP/main.m:185-187, clutter added as
    MP/fun_add_clutter.m:6-9, per-segment SCR gains in the manner of MP/fun_SCR.m:23,31-36
    (power taken as |.|^2; the reference's complex ``.^2`` is an input-scaling quirk, not path maths)."""
    rng = np.random.default_rng(seed)
    clutter = np.rint(rng.normal(0.0, 150.0, size=(P, R))) + 1j * np.rint(rng.normal(0.0, 150.0, size=(P, R)))
    pulse1 = mcode.pulse1_mp()
    pulse2, pulse3 = mcode.load_pulse_literals()
    fc = 5500e6
    prt = 64.88e-6
    lam = C_LIGHT / fc
    fd = 2.0 * velocity / lam
    ph = np.exp(2j * np.pi * fd * prt * np.arange(P))
    cell = mcode.mround(range_m / 6.0)
    segs = ((0, 82, pulse1, scr_db + 10.0), (82, 242, pulse2, scr_db), (324, R - 324, pulse3, scr_db))
    echo = np.zeros((P, R), dtype=np.complex128)
    for start, length, pulse, scr in segs:
        s = np.zeros(length, dtype=np.complex128)
        n = min(len(pulse), length - cell)
        s[cell:cell + n] = pulse[:n]
        ps = np.mean(np.abs(s) ** 2) + np.finfo(float).eps
        for i in range(P):
            pe = np.mean(np.abs(clutter[i, start:start + length]) ** 2)
            g = pe * 10.0 ** (scr / 10.0) / ps
            echo[i, start:start + length] = s * np.sqrt(g) * ph[i]
    return echo + clutter


S2_CFAR = dict(refR=5, saveR=7, T_R=5.0, methR=0, refV=2, saveV=1, T_V=5.0, methV=0, n0=0, rflag=1)
S3_CFAR = dict(refR=5, saveR=7, T_R=5.0, methR=0, refV=5, saveV=7, T_V=5.0, methV=0, n0=0, rflag=1)
S5_CFAR = dict(refR=5, saveR=7, T_R=7.0, methR=0, refV=5, saveV=7, T_V=7.0, methV=0, n0=0, rflag=1)
S1_CFAR = dict(refR=5, saveR=7, T_R=5.0, methR=0, refV=5, saveV=7, T_V=5.0, methV=0, n0=10, rflag=1)


def cfar_tuple(d):
    return (d["refR"], d["saveR"], d["T_R"], d["methR"], d["refV"], d["saveV"], d["T_V"], d["methV"],
            d["n0"], d["rflag"])


def to_dbf24(lanes, channel_num):
    """Encode integer-valued complex lanes [prt][range][col] as DBF-type PRT payloads (uint8 [prt][bytes]) the way
    FrameDataRead_xzr.m:111-119,130-135,163 expects them: rows of 6*channel_num + pad bytes, 3-byte little-endian
    two's-complement words I0 Q0 I1 Q1 ..., each PRT padded to a multiple of 64 bytes."""
    from . import mcode
    lanes = np.asarray(lanes)
    n_prt, n, ncol = lanes.shape
    sig, pad, osp = mcode.dbf24_payload_size(n, channel_num)
    W = channel_num * 6 + osp
    assert ncol * 6 <= W
    out = np.zeros((n_prt, sig + pad), dtype=np.uint8)
    rows = out[:, :sig].reshape(n_prt, n, W)
    for col in range(ncol):
        for part, val in ((0, lanes[:, :, col].real), (1, lanes[:, :, col].imag)):
            w = np.round(val).astype(np.int64) & 0xFFFFFF
            o = col * 6 + part * 3
            rows[:, :, o] = w & 0xFF
            rows[:, :, o + 1] = (w >> 8) & 0xFF
            rows[:, :, o + 2] = (w >> 16) & 0xFF
    return out


def frame_prt_dbf24(payload_bytes, n_range, frame_no=0, prt_no=0, channel_num=13, servo=0):
    """Wrap one DBF-type PRT payload (already padded, see to_dbf24) in the 64 B head / 128 B realtime / 64 B tail framing
    with data_type = 2 (FrameDataRead_xzr.m:62-119,184)."""
    head = np.zeros(16, dtype="<u4")
    head[0] = frame_no
    head[2] = prt_no & 0xFFFF
    head[3] = channel_num & 0xFF
    head[4] = servo & 0xFFFF
    head[6] = n_range
    head[7] = 2
    return head.tobytes() + bytes(128) + bytes(np.ascontiguousarray(payload_bytes, dtype=np.uint8)) + bytes(64)
