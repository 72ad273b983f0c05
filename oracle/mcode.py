"""Loop-faithful double-precision transcription of the reference M functions (TEST ORACLE).

Every function keeps the reference's loop structure and 1-based index arithmetic (indices are
converted to 0-based only at the moment of array access) so that it can be read side by side with
the M file it cites.  Paths are relative to the reference root; ``MP`` = ``MatlabProcess_xuzerui``,
``CW`` = ``MatlabProcess_xuzerui/CFAR_WangCai``.

The arithmetic the M-code delegates to MATLAB built-ins (``fft ifft filter kaiser fftshift mean max
find round circshift grpdelay``; MATLAB R2025a + Signal Processing Toolbox, closed source, not under
/root/reference) is restated from the published definitions of those functions.
"""
import math
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, "tests", "golden")


class MatlabError(Exception):
    """An error the MATLAB interpreter would raise for the same call (identifier in ``ident``)."""

    def __init__(self, ident, msg):
        super().__init__("%s: %s" % (ident, msg))
        self.ident = ident


# ----------------------------------------------------------------------------------------------
# MATLAB built-in restatements
# ----------------------------------------------------------------------------------------------
def mround(x):
    """MATLAB ``round``: half away from zero (Python's ``round`` is half-to-even)."""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def kaiser(n, beta):
    """Signal Processing Toolbox ``kaiser(n,beta)`` as a length-n column (1-D array).

    w[k] = I0(beta*sqrt(1-((k-(n-1)/2)/((n-1)/2))^2)) / I0(beta), k = 0..n-1; kaiser(1,b) = 1.
    Pinned against MP/kaiser_win.mat by tests/test_oracle_golden.py.
    """
    n = int(n)
    if n == 1:
        return np.ones(1)
    k = np.arange(n, dtype=np.float64)
    alpha = (n - 1) / 2.0
    r = (k - alpha) / alpha
    return np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / np.i0(beta)


def matlab_filter_fir(b, x):
    """``filter(b,1,x)`` for a vector x: y[n] = sum_k b[k] x[n-k], same length as x."""
    b = np.asarray(b)
    x = np.asarray(x)
    y = np.zeros(x.shape[0], dtype=np.result_type(b.dtype, x.dtype, np.float64))
    for n in range(x.shape[0]):
        acc = 0.0
        for k in range(min(n + 1, b.shape[0])):
            acc = acc + b[k] * x[n - k]
        y[n] = acc
    return y


def grpdelay_mean_round(b):
    """``round(mean(grpdelay(b)))`` for a real FIR b (MTD/fun_lss_pulse_compression.m:47).

    grpdelay evaluates -d(phase)/dw at 512 points on [0,pi); computed here from its definition
    gd(w) = Re{ sum k b_k e^{-jwk} / sum b_k e^{-jwk} }, singular points (|B|~0) replaced by 0 as
    the toolbox does.
    """
    b = np.asarray(b, dtype=np.float64)
    k = np.arange(b.size)
    w = np.pi * np.arange(512) / 512.0
    e = np.exp(-1j * np.outer(w, k))
    num = e @ (k * b)
    den = e @ b
    small = np.abs(den) < 10 * np.finfo(float).eps * max(1.0, np.abs(b).sum())
    gd = np.real(num / np.where(small, 1.0, den))
    gd[small] = 0.0
    return mround(float(np.mean(gd)))


# ----------------------------------------------------------------------------------------------
# fixtures: literal constants of the reference (inputs, not outputs)
# ----------------------------------------------------------------------------------------------
FILTER_COEF_INT = np.array(
    [-9, -7, -2, 10, 27, 40, 42, 24, -13, -57, -89, -86, -30, 77, 220, 364, 471, 511, 471, 364, 220, 77,
     -30, -86, -89, -57, -13, 24, 42, 40, 27, 10, -2, -7, -9], dtype=np.float64)
"""MP/fun_lss_pulse_compression.m:21 (identical at MTD/fun_lss_pulse_compression.m:31)."""


def load_pulse_literals():
    """pulse2 (75) and pulse3 (160): MP/fun_MTD_produce.m:54-60 (tests/golden/pulse_literals.npz)."""
    z = np.load(os.path.join(_GOLDEN, "pulse_literals.npz"))
    assert np.array_equal(z["filter_coef"], FILTER_COEF_INT)
    return z["pulse2"], z["pulse3"]


def load_ref(name):
    """``refData`` of MP/refDDCDataMF1.mat or MP/refDBFDataMF1.mat (67 taps)."""
    return np.load(os.path.join(_GOLDEN, name + ".npy"))


def pulse1_mp():
    """MP/fun_MTD_produce.m:24-27,47,51: ``sin(2*pi*t1+pi/2)``, t1=-tao1/2:ts:tao1/2-ts (7 points)."""
    fs = 25e6
    ts = 1 / fs
    tao1 = 0.28e-6
    return np.sin(2 * np.pi * matlab_colon(-tao1 / 2, ts, tao1 / 2 - ts) + np.pi / 2)


def matlab_colon(a, d, b):
    """``a:d:b`` as MATLAB builds it (MathWorks' published COLONOP replication): the element count tolerates rounding of the end
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


# ----------------------------------------------------------------------------------------------
# A1  DDC unpack  (FrameDataRead_xzr.m:104-119,138,150-156)
# ----------------------------------------------------------------------------------------------
def ddc_payload_size(pulse_data_num, channel_num):
    """(sig_data_size, pad_size) in bytes for data_type==1.  FrameDataRead_xzr.m:109,115-119."""
    sig = pulse_data_num * channel_num * 2 * 2
    pad = (64 - sig % 64) if sig % 64 > 0 else 0
    return sig, pad


def unpack_ddc_i16(signal_data_raw, pulse_data_num, channel_num):
    """One PRT payload (uint8 bytes incl. padding) -> (pulse_data_num, channel_num) complex128."""
    raw = np.frombuffer(np.ascontiguousarray(signal_data_raw, dtype=np.uint8).tobytes(), dtype="<i2")  # :138 typecast
    sig = raw[: pulse_data_num * channel_num * 2]                        # :150
    sig = sig.reshape(-1, channel_num * 2)                               # :151 reshape(ch*2,[]).'
    sig = sig.astype(np.float64)                                         # :152
    I = sig[:, 0::2]                                                     # :154 cols 1:2:end
    Q = sig[:, 1::2]                                                     # :155 cols 2:2:end
    return I + 1j * Q                                                    # :156


def dbf24_payload_size(pulse_data_num, channel_num):
    """(sig_data_size, pad_size, one_sample_pad) in bytes for data_type==2.  FrameDataRead_xzr.m:111-119."""
    one_sample_pad = 8 - (6 * channel_num) % 8                              # :111 (8, not 0, when already aligned)
    sig = pulse_data_num * channel_num * 2 * 3 + pulse_data_num * one_sample_pad
    pad = (64 - sig % 64) if sig % 64 > 0 else 0
    return sig, pad, one_sample_pad


def unpack_dbf24(signal_data_raw, pulse_data_num, channel_num):
    """One DBF-type PRT payload (uint8 bytes incl. padding) -> (pulse_data_num, ncol) complex128.

    FrameDataRead_xzr.m:130-135,163: rows of ``channel_num*6 + one_sample_pad`` bytes, 3-byte little-endian
    words, ``> 2^23`` (not ``>=``) mapped to negative, I = cols 1:2:end, Q = cols 2:2:end.
    DEVIATION, stated: the M-code does this arithmetic on ``uint8`` data (read_continuous_file_stream.m:89 reads
    ``*uint8``), which saturates at 255 -- the source comment at :130 says the block "has problems, not fixed
    yet".  The restatement widens the bytes to double first, i.e. implements the 24-bit format the comment
    describes, and keeps the ``> 2^23`` quirk (0x800000 decodes to +8388608).
    """
    sig, pad, osp = dbf24_payload_size(pulse_data_num, channel_num)
    raw = np.ascontiguousarray(signal_data_raw, dtype=np.uint8).ravel()
    W = channel_num * 2 * 3 + osp
    data_temp = raw[: raw.size - pad].astype(np.float64).reshape(-1, W)      # :132 reshape(W,[]).'
    c1 = np.arange(1, W - 3 + 1, 3)                                         # 1:3:end-3
    c2 = np.arange(2, W - 2 + 1, 3)                                         # 2:3:end-2
    c3 = np.arange(3, W + 1, 3)                                             # 3:3:end
    if not (c1.size == c2.size == c3.size):
        raise MatlabError("MATLAB:sizeDimensionsMustMatch", "Arrays have incompatible sizes for this operation")
    parsed = data_temp[:, c1 - 1] + data_temp[:, c2 - 1] * 2 ** 8 + data_temp[:, c3 - 1] * 2 ** 16     # :133
    parsed[parsed > 2 ** 23] -= 2 ** 24                                     # :134-135
    if parsed.shape[1] % 2:
        raise MatlabError("MATLAB:sizeDimensionsMustMatch", "Arrays have incompatible sizes for this operation")
    return parsed[:, 0::2] + 1j * parsed[:, 1::2]                           # :163


# ----------------------------------------------------------------------------------------------
# A4  fun_pulse_compression  (MP/fun_pulse_compression.m:1-24)
# ----------------------------------------------------------------------------------------------
def fun_pulse_compression(s0, s_echo):
    s0 = np.atleast_1d(np.asarray(s0)).ravel()
    s_echo = np.atleast_1d(np.asarray(s_echo)).ravel()
    h = np.conj(s0[::-1])                                                # :4
    point_pulse = h.shape[0]                                             # :5-6
    point_prt = s_echo.shape[0]                                          # :14-15
    n = point_pulse + point_prt - 1                                      # :16
    if n < 1:
        return np.zeros(0, dtype=np.complex128)
    S = np.fft.fft(s_echo, n)                                            # :19
    H = np.fft.fft(h, n)                                                 # :20
    Y = S * H                                                            # :21
    return np.fft.ifft(Y, n)                                             # :22


# ----------------------------------------------------------------------------------------------
# A3  fun_lss_pulse_compression  (5-arg MP/ and 9-arg MTD/)
# ----------------------------------------------------------------------------------------------
def fun_lss_pulse_compression_mp(echo, pulse1, pulse2, pulse3):
    """MP/fun_lss_pulse_compression.m:3-45 (show_PC = 0).  ``pulse1`` is accepted and unused."""
    echo = np.asarray(echo)
    p1, p2, p3 = 82, 242, 707                                            # :6-8 (p3 unused by the code)
    m, n = echo.shape                                                    # :11
    if n < p1 + p2:
        raise MatlabError("MATLAB:badsubscript", "Index in position 2 exceeds array bounds")
    s1 = echo[:, 0:p1]                                                   # :14
    s2 = echo[:, p1:p1 + p2]                                             # :15
    s3 = echo[:, p1 + p2:n]                                              # :16
    out = np.zeros((m, n), dtype=np.complex128)                          # :18
    b = FILTER_COEF_INT / FILTER_COEF_INT.max()                          # :21-22
    for i in range(m):                                                   # :24
        pc1 = matlab_filter_fir(b, s1[i, :]) / 1.2                       # :25-26
        pc2 = fun_pulse_compression(pulse2, s2[i, :])                    # :28
        pc3 = fun_pulse_compression(pulse3, s3[i, :])                    # :29
        out[i, 0:p1] = pc1[0:p1]                                         # :31
        r2 = pc2[74:]                                                    # :36 (75:end)
        if r2.shape[0] != p2:
            raise MatlabError("MATLAB:subsassigndimmismatch", "pulse2 must have 75 samples")
        out[i, p1:p1 + p2] = r2
        r3 = pc3[159:]                                                   # :37 (160:end)
        if r3.shape[0] != n - p1 - p2:
            raise MatlabError("MATLAB:subsassigndimmismatch", "pulse3 must have 160 samples")
        out[i, p1 + p2:n] = r3
    return out


def fun_lss_pulse_compression_mtd(echo, pulse1, pulse2, pulse3, p1, p2, p3):
    """MTD/fun_lss_pulse_compression.m:17-80 (show_PC = 0; ``params`` only feeds the plot)."""
    echo = np.asarray(echo)
    m, n = echo.shape                                                    # :20
    p1, p2, p3 = int(p1), int(p2), int(p3)
    if n < p1 + p2:
        raise MatlabError("MATLAB:badsubscript", "Index in position 2 exceeds array bounds")
    s1 = echo[:, 0:p1]                                                   # :23
    s2 = echo[:, p1:p1 + p2]                                             # :24
    s3 = echo[:, p1 + p2:n]                                              # :25
    out = np.zeros((m, n), dtype=np.complex128)                          # :27
    b = FILTER_COEF_INT / FILTER_COEF_INT.max()                          # :31-32
    pulse2 = np.atleast_1d(pulse2).ravel()
    pulse3 = np.atleast_1d(pulse3).ravel()
    for i in range(m):                                                   # :36
        pc1 = matlab_filter_fir(b, s1[i, :]) / 1.2                       # :38-39
        pc2 = fun_pulse_compression(pulse2, s2[i, :])                    # :41
        pc3 = fun_pulse_compression(pulse3, s3[i, :])                    # :42
        delay1 = grpdelay_mean_round(b)                                  # :47
        tmp = np.roll(pc1, -delay1)                                      # :50 circshift(x,-d)
        out[i, 0:p1] = tmp[0:p1]                                         # :51
        o2 = pulse2.shape[0]                                             # :58
        if o2 + p2 - 1 > pc2.shape[0]:
            raise MatlabError("MATLAB:badsubscript", "Index exceeds the number of array elements")
        out[i, p1:p1 + p2] = pc2[o2 - 1:o2 + p2 - 1]                     # :60
        o3 = pulse3.shape[0]                                             # :63
        if o3 + p3 - 1 > pc3.shape[0]:
            raise MatlabError("MATLAB:badsubscript", "Index exceeds the number of array elements")
        if p1 + p2 + p3 > n:
            # assignment past the last column would grow the matrix in MATLAB; by the check above
            # p3 <= n-p1-p2 always holds here, so this cannot trigger -- kept for clarity.
            raise MatlabError("MATLAB:badsubscript", "segment sizes exceed the PRT length")
        out[i, p1 + p2:p1 + p2 + p3] = pc3[o3 - 1:o3 + p3 - 1]           # :65
    return out


# ----------------------------------------------------------------------------------------------
# A5  fun_Process_MTD  (MP/fun_Process_MTD.m:3-33)
# ----------------------------------------------------------------------------------------------
def matlab_fftshift(v):
    n = v.shape[0]
    return np.roll(v, n // 2)          # fftshift moves element ceil(n/2)+1 (1-based) to the front


def fun_Process_MTD(ProSignal, Len_PRT, Num_PRTperFrame, beta=8.0, window=None):
    ProSignal = np.asarray(ProSignal)
    Len_PRT = int(Len_PRT)
    P = int(Num_PRTperFrame)
    if ProSignal.shape[0] != P and ProSignal.shape[0] != 1 and P != 1:
        raise MatlabError("MATLAB:sizeDimensionsMustMatch", "Arrays have incompatible sizes")
    if Len_PRT > ProSignal.shape[1]:
        raise MatlabError("MATLAB:badsubscript", "Index in position 2 exceeds array bounds")
    w = kaiser(P, beta) if window is None else np.asarray(window, dtype=np.float64)   # :13-14
    out = np.zeros((P, Len_PRT))                                         # :17
    for idx in range(Len_PRT):                                           # :20
        sw = ProSignal[:, idx] * w                                       # :22
        f = matlab_fftshift(np.fft.fft(sw, P))                           # :24
        out[:, idx] = np.abs(f)                                          # :26,29
    return out


# ----------------------------------------------------------------------------------------------
# A6  fun_0v_pressing  (MP/ & MTD/: div=150; CW/: div=20)
# ----------------------------------------------------------------------------------------------
def zero_v_rows(prtNum, div):
    """1-based inclusive row range zeroed by fun_0v_pressing (MP/fun_0v_pressing.m:4,6)."""
    z = mround(prtNum / 2)
    h = mround(prtNum / div)
    return z - h, z + h


def fun_0v_pressing(MTD, div=150):
    MTD = np.array(MTD, copy=True)
    lo, hi = zero_v_rows(MTD.shape[0], div)
    if lo < 1 or hi > MTD.shape[0]:
        raise MatlabError("MATLAB:badsubscript", "Index in position 1 is invalid")
    MTD[lo - 1:hi, :] = 0
    return MTD


# ----------------------------------------------------------------------------------------------
# A11/A12 optional pre-stages
# ----------------------------------------------------------------------------------------------
def fun_iSTC(echo, stc_ini):
    """MP/fun_iSTC.m:2-17 with the (missing) STC curve file content passed in as ``stc_ini`` (dB)."""
    echo = np.asarray(echo)
    m, n = echo.shape
    stc = np.zeros(n)
    stc_ini = np.asarray(stc_ini, dtype=np.float64).ravel()
    if stc_ini.size > n:
        stc = np.zeros(stc_ini.size)      # MATLAB grows stc; the .* below then errors
        raise MatlabError("MATLAB:sizeDimensionsMustMatch", "STC curve longer than the PRT")
    stc[: stc_ini.size] = stc_ini                                         # :8-9
    out = np.zeros((m, n), dtype=np.complex128)
    g = 10.0 ** (stc / 20.0)
    for i in range(m):                                                   # :12
        out[i, :] = echo[i, :] * g                                       # :14
    return stc, out


def fun_Process_MTI(ProSignal, lag=30):
    """MP/fun_Process_MTI.m:1-28 (the running mean at :10-14 is computed and unused)."""
    ProSignal = np.asarray(ProSignal)
    P, R = ProSignal.shape
    out = np.zeros((P, R), dtype=ProSignal.dtype)
    for i in range(1, P - lag + 1):                                      # :20
        out[i - 1, :] = ProSignal[i + lag - 1, :] - ProSignal[i - 1, :]  # :21
    return out


# ----------------------------------------------------------------------------------------------
# A2  fun_MTD_produce
# ----------------------------------------------------------------------------------------------
def fun_MTD_produce_mp(echo):
    """MP/fun_MTD_produce.m:3-79 (1-arg canonical API; plotting off)."""
    pulse1 = pulse1_mp()                                                 # :51
    pulse2, pulse3 = load_pulse_literals()                               # :54-60
    Echo_0 = fun_lss_pulse_compression_mp(echo, pulse1, pulse2, pulse3)  # :67
    m, n = Echo_0.shape                                                  # :77
    MTD_Signal = fun_Process_MTD(Echo_0, n, m)                           # :78
    return fun_0v_pressing(MTD_Signal, 150)                              # :79


def ideal_pulses_mtd(params):
    """MTD/fun_MTD_produce.m:37-38,45-51,61-69: ideal simple pulse + two LFM chirps."""
    fs = float(params["fs"])
    ts = 1 / fs
    B = float(params["B"])
    tao1, tao2, tao3 = [float(t) for t in params["tao"][:3]]
    K2 = -B / tao2
    K3 = B / tao3
    t1 = matlab_colon(-tao1 / 2, ts, tao1 / 2 - ts)
    t2 = matlab_colon(-tao2 / 2, ts, tao2 / 2 - ts)
    t3 = matlab_colon(-tao3 / 2, ts, tao3 / 2 - ts)
    pulse1 = np.sin(2 * np.pi * t1 + np.pi / 2)
    pulse2 = np.exp(1j * 2 * np.pi * (0.5 * K2 * t2 ** 2))
    pulse3 = np.exp(1j * 2 * np.pi * (0.5 * K3 * t3 ** 2))
    return pulse1, pulse2, pulse3


def fun_MTD_produce_mtd(echo, params):
    """MTD/fun_MTD_produce.m:12-102 (2-arg API; debug plots off)."""
    pulse1, pulse2, pulse3 = ideal_pulses_mtd(params)
    pp = params["point_prt"]
    Echo_0 = fun_lss_pulse_compression_mtd(echo, pulse1, pulse2, pulse3, pp[1], pp[2], pp[3])   # :86
    m, n = Echo_0.shape                                                  # :97
    MTD_Signal = fun_Process_MTD(Echo_0, n, m)                           # :98
    return fun_0v_pressing(MTD_Signal, 150)                              # :102


# ----------------------------------------------------------------------------------------------
# A8/A9  1-D CA-CFAR  (CW/Function_CFAR1D_sub.m, CW/Function_CFAR1D_sub_fixCells.m)
# ----------------------------------------------------------------------------------------------
def _mean_cols(data, rows, c1, c2):
    """``mean(data(rows, c1:c2), 2)`` with 1-based inclusive columns; sequential left-to-right sum."""
    ncol = data.shape[1]
    if c1 < 1 or c2 > ncol:
        raise MatlabError("MATLAB:badsubscript", "Index in position 2 exceeds array bounds")
    acc = np.zeros(len(rows))
    for c in range(c1, c2 + 1):
        acc = acc + data[rows, c - 1]
    return acc / (c2 - c1 + 1)


def Function_CFAR1D_sub(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod):
    data = np.asarray(datamatrix, dtype=np.float64)
    nrow, ncol = data.shape                                              # :14
    out = np.zeros((nrow, ncol))                                         # :15
    rows = np.arange(nrow)
    for y in range(1, ncol + 1):                                         # :17
        refL1 = y - (saveCellNum + refCellNum)                           # :25
        refL2 = y - saveCellNum - 1                                      # :26
        refR1 = y + saveCellNum + 1                                      # :27
        refR2 = y + saveCellNum + refCellNum                             # :28
        if refL1 >= 1:                                                   # :30
            refL = _mean_cols(data, rows, refL1, refL2)
        else:
            refL = _mean_cols(data, rows, refR1, refR2)                  # :33
        if refR2 <= ncol:                                                # :35
            refR = _mean_cols(data, rows, refR1, refR2)
        else:
            refR = _mean_cols(data, rows, refL1, refL2)                  # :38
        used = np.maximum(refL, refR) if CFARmethod == 0 else np.minimum(refL, refR)   # :40-44
        thr = used * T_CFAR                                              # :45
        out[:, y - 1] = (data[:, y - 1] >= thr).astype(np.float64)       # :46,68
    return out


def Function_CFAR1D_sub_fixCells(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod, rowCellsFix, colCellsFix):
    data = np.asarray(datamatrix, dtype=np.float64)
    if data.ndim == 1:
        data = data[None, :]
    nrow, ncol = data.shape                                              # :17
    out = np.zeros((nrow, ncol))                                         # :18
    rows = np.atleast_1d(np.asarray(rowCellsFix, dtype=np.int64)) - 1
    for y in np.atleast_1d(np.asarray(colCellsFix, dtype=np.int64)):     # :23-24
        y = int(y)
        refL1 = y - (saveCellNum + refCellNum)                           # :34
        refL2 = y - saveCellNum - 1
        refR1 = y + saveCellNum + 1
        refR2 = y + saveCellNum + refCellNum                             # :37
        if refL1 >= 1:                                                   # :39
            refL = _mean_cols(data, rows, refL1, refL2)
        else:
            refL = _mean_cols(data, rows, refR1, refR2)
        if refR2 <= ncol:                                                # :44
            refR = _mean_cols(data, rows, refR1, refR2)
        else:
            refR = _mean_cols(data, rows, refL1, refL2)
        used = np.maximum(refL, refR) if CFARmethod == 0 else np.minimum(refL, refR)   # :50-54
        thr = used * T_CFAR                                              # :55
        out[rows, y - 1] = (data[rows, y - 1] >= thr).astype(np.float64)  # :58,85
    return out


# ----------------------------------------------------------------------------------------------
# A7  executeCFAR  (CW/executeCFAR.m:1-92)
# ----------------------------------------------------------------------------------------------
def executeCFAR(mtd, refCells_R, saveCells_R, T_CFAR_R, CFARmethod_R,
                refCells_V, saveCells_V, T_CFAR_V, CFARmethod_V, MTD_0_num, rCFARDetect_Flag):
    mtd = np.asarray(mtd, dtype=np.float64)
    vorg, rorg = mtd.shape                                               # :21
    n0 = int(MTD_0_num)
    lo, hi = n0 + 2, vorg - n0                                           # :23 rows lo:hi (1-based)
    if lo < 1 or hi > vorg:
        raise MatlabError("MATLAB:badsubscript", "Index in position 1 exceeds array bounds")
    used = mtd[lo - 1:hi, :]
    vnum, rnum = used.shape                                              # :24
    resV = Function_CFAR1D_sub(used.T, refCells_V, saveCells_V, T_CFAR_V, CFARmethod_V).T    # :28
    flagV = np.zeros((vorg, rorg))                                       # :30
    flagV[lo - 1:hi, :] = resV                                           # :31
    flag = np.zeros((vorg, rorg))                                        # :33
    if rCFARDetect_Flag:                                                 # :35
        # find() returns column-major order                             # :36
        cc, rr = np.nonzero(resV.T)
        vCell_row = rr + 1
        rCell_col = cc + 1
        resR = np.zeros((vnum, rnum))                                    # :78 / :86
        rows1, cols2 = [], []
        for mm in range(vCell_row.shape[0]):                             # :45
            v = int(vCell_row[mm])
            r = int(rCell_col[mm])
            cells = [r - 1, r, r + 1]                                    # :38,50
            cells = [c for c in cells if 0 < c <= rnum]                  # :52-57
            dataUsedTemp = used[v - 1:v, :]                              # :59
            det = Function_CFAR1D_sub_fixCells(dataUsedTemp, refCells_R, saveCells_R, T_CFAR_R,
                                               CFARmethod_R, 1, cells)    # :61
            nz = np.nonzero(det[0])[0] + 1                               # :64
            if nz.size > 0:                                              # :65
                rows1.append(v)
                if nz.size > 1:                                          # :68
                    I = int(np.argmax(dataUsedTemp[0, nz - 1]))          # :69 first max
                    cols2.append(int(nz[I]))                             # :70
                else:
                    cols2.append(int(nz[0]))                             # :72
        for a, b in zip(rows1, cols2):                                   # :81-83
            resR[a - 1, b - 1] = 1
        flag[lo - 1:hi, :] = resR                                        # :89
    else:
        flag = flagV.copy()                                              # :91
    return flag, flagV


def fun_CFARflag(MTD_data, refCells_R, saveCells_R, T_CFAR_R, CFARmethod_R, refCells_V, saveCells_V,
                 T_CFAR_V, CFARmethod_V, MTD_0_num, rCFARDetect_Flag, segments=((1, 82), (83, 318), (319, 868))):
    """CW/main_cfar.m:142-161 (local function; stays M-code in the drop-in, restated for tests)."""
    MTD_data = np.asarray(MTD_data, dtype=np.float64)
    out = np.zeros(MTD_data.shape)                                       # :156
    for a, b in segments:                                                # :143-145
        f, _ = executeCFAR(MTD_data[:, a - 1:b], refCells_R, saveCells_R, T_CFAR_R, CFARmethod_R,
                           refCells_V, saveCells_V, T_CFAR_V, CFARmethod_V, MTD_0_num, rCFARDetect_Flag)
        out[:, a - 1:b] = f                                              # :157-159
    return out


# ----------------------------------------------------------------------------------------------
# f4  DMX script variant  (CW/DMX_SignalProcessing_main_xzr.m) -- one frame, two monopulse beams
# ----------------------------------------------------------------------------------------------
def hamming(n):
    """Signal Processing Toolbox ``hamming(n)`` (symmetric): 0.54 - 0.46*cos(2*pi*k/(n-1)); hamming(1) = 1."""
    n = int(n)
    if n == 1:
        return np.ones(1)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))


def dmx_match_filter(ref, beta=4.5, power_norm=True):
    """matchWaveform2 .* mfWh.' of CW/DMX_SignalProcessing_main_xzr.m:158-166,187-202 (winType 3 = kaiser(len,4.5))."""
    w = np.asarray(ref, dtype=np.complex128).ravel()                     # :160 refData.'
    if power_norm:
        w = w / np.sqrt(np.sum(np.abs(w) ** 2))                          # :166 norm()
    return w * kaiser(w.size, beta)                                      # :189,202


def dmx_frame(left, right, point_short, filter_coef, match_windowed, FFT_num, mtdWh, mtd_FFT_num, MTD_0_num):
    """CW/DMX_SignalProcessing_main_xzr.m:332-353 (split + pulse compression), :414-426 (MTD, sum / difference),
    :462-465 (zero-Doppler blanking of the sums).  Returns (sum_short, diff_short, sum_long, diff_long)."""
    left = np.asarray(left, dtype=np.complex128)
    right = np.asarray(right, dtype=np.complex128)
    prtNum = left.shape[0]
    matchF2 = np.conj(np.fft.fft(match_windowed, FFT_num))               # :202
    mtdWh = np.asarray(mtdWh, dtype=np.float64).reshape(prtNum, 1)
    mags = []
    for beam in (left, right):
        short = beam[:, :point_short]                                    # :332,335
        long_ = beam[:, point_short:]                                    # :333,336
        mf_short = np.zeros(short.shape, dtype=np.complex128)
        for i in range(prtNum):                                          # :344-345 filter along range, per PRT
            mf_short[i, :] = matlab_filter_fir(filter_coef, short[i, :]) if point_short else short[i, :]
        spec = np.fft.fft(long_, FFT_num, axis=1)                        # :348-349 (zero-padded to FFT_num)
        mf_long = np.fft.ifft(spec * matchF2[None, :], axis=1)           # :352-353
        mtd_short = np.fft.fft(mf_short * mtdWh, mtd_FFT_num, axis=0)    # :414-415
        mtd_long = np.fft.fft(mf_long * mtdWh, mtd_FFT_num, axis=0)      # :417-418
        mags.append((np.abs(mtd_short), np.abs(mtd_long)))
    sum_short = mags[0][0] + mags[1][0]                                  # :421
    sum_long = mags[0][1] + mags[1][1]                                   # :422
    diff_short = mags[1][0] - mags[0][0]                                 # :425
    diff_long = mags[1][1] - mags[0][1]                                  # :426
    if MTD_0_num is not None and MTD_0_num >= 0:
        rows = list(range(0, MTD_0_num + 1)) + list(range(mtd_FFT_num - MTD_0_num, mtd_FFT_num))   # :463 (1-based there)
        sum_short[rows, :] = 0                                           # :464
        sum_long[rows, :] = 0                                            # :465
    return sum_short, diff_short, sum_long, diff_long


# ----------------------------------------------------------------------------------------------
# f3  motionParaMeasure  (CW/motionParaMeasure.m:1-88) -- post-CFAR range / velocity / elevation measurement
# ----------------------------------------------------------------------------------------------
def matlab_spline_eval(y, xq):
    """``interp1(0:n-1, y, xq, 'spline')``: not-a-knot cubic spline (MATLAB ``spline``); n = 2 -> line, n = 3 -> parabola."""
    from scipy.interpolate import CubicSpline
    y = np.asarray(y, dtype=np.float64)
    n = y.size
    x = np.arange(n, dtype=np.float64)
    if n == 1:
        return np.full(np.shape(xq), y[0])
    if n == 2:
        return y[0] + (y[1] - y[0]) * np.asarray(xq)
    return CubicSpline(x, y, bc_type="not-a-knot")(np.asarray(xq))


def motionParaMeasure(mtd_sum, mtd_diff, flags, extraDots, rScale, deltaR, rInterpTimes, vScale, deltaV, vInterpTimes,
                      kValues, beamPosNum, beamAngleStep, freInd, eleAngleComp, eleAngleSysErr, MTD_0_num):
    mtd_sum = np.asarray(mtd_sum, dtype=np.float64)
    mtd_diff = np.asarray(mtd_diff, dtype=np.float64)
    flags = np.asarray(flags)
    vCellNum, rCellNum = flags.shape                                      # :5
    cc, rr = np.nonzero(flags.T)                                          # :6 find(): column-major
    vInd, rInd = rr + 1, cc + 1
    k = int(extraDots)
    ext = np.arange(-k, k + 1)                                            # :8
    rEst, vEst, eEst = [], [], []
    for mm in range(vInd.size):                                           # :17
        v, r = int(vInd[mm]), int(rInd[mm])
        cells = ext + r                                                   # :22
        if cells.min() <= 0:                                              # :24-27
            cells = 1 + np.arange(0, 2 * k + 1)
        if cells.max() > rCellNum:                                        # :29-32
            cells = rCellNum - np.arange(0, 2 * k + 1)
        cells = np.sort(cells)                                            # :33
        if cells.min() < 1 or cells.max() > rCellNum:
            raise MatlabError("MATLAB:badsubscript", "Index exceeds array bounds")
        data = mtd_sum[v - 1, cells - 1]                                  # :36
        nq = int(math.floor((cells[-1] - cells[0]) * rInterpTimes + 1e-9)) + 1
        xq = cells[0] + np.arange(nq) / float(rInterpTimes)               # :37
        dq = matlab_spline_eval(data, xq - cells[0])                      # :38
        rCellMax = xq[int(np.argmax(dq))]                                 # :39-42
        r_est = rScale[r - 1] + (rCellMax - r) * deltaR                   # :43
        cells = ext + v                                                   # :49
        if cells.min() <= MTD_0_num + 1:                                  # :51-54
            cells = (MTD_0_num + 2) + np.arange(0, 2 * k + 1)
        if cells.max() > vCellNum - MTD_0_num:                            # :56-59
            cells = (vCellNum - MTD_0_num) - np.arange(0, 2 * k + 1)
        cells = np.sort(cells)                                            # :60
        if cells.min() < 1 or cells.max() > vCellNum:
            raise MatlabError("MATLAB:badsubscript", "Index exceeds array bounds")
        data = mtd_sum[cells - 1, r - 1]                                  # :63
        nq = int(math.floor((cells[-1] - cells[0]) * vInterpTimes + 1e-9)) + 1
        xq = cells[0] + np.arange(nq) / float(vInterpTimes)               # :64
        dq = matlab_spline_eval(data, xq - cells[0])                      # :65
        vCellMax = xq[int(np.argmax(dq))]                                 # :66-69
        fx = int(np.fix(vCellMax))
        v_est = vScale[fx - 1] - (vCellMax - fx) * deltaV                 # :70
        ratio = mtd_diff[v - 1, r - 1] / mtd_sum[v - 1, r - 1]            # :76-78
        e_est = beamPosNum * beamAngleStep + 2.5 - ratio * kValues[int(freInd), int(beamPosNum)] + eleAngleComp + eleAngleSysErr   # :79
        rEst.append(r_est)
        vEst.append(v_est)
        eEst.append(e_est)
    return np.array(rEst), np.array(vEst), np.array(eEst)


# ----------------------------------------------------------------------------------------------
# f2  read_continuous_file_stream.m:22-168 + DDC branch of FrameDataRead_xzr.m:57-198 (TEST ORACLE)
# ----------------------------------------------------------------------------------------------
class ContinuousFileStream:
    """The persistent state machine of read_continuous_file_stream.m, file names per DataFullPathGen.m:10-16."""

    def __init__(self, directory):
        self.dir = directory
        self.is_open = False
        self.data = b""
        self.pos = 0
        self.max_len = 0
        self.index = 0                                                     # :43

    def _name(self, i):
        name = "1.00000%d.bin" % i if i < 10 else ("1.0000%d.bin" % i if i < 100 else "1.000%d.bin" % i)
        return os.path.join(self.dir, name)

    def _open(self, i):
        try:
            self.data = open(self._name(i), "rb").read()
        except OSError:
            return False
        self.max_len, self.pos, self.is_open = len(self.data), 0, True
        return True

    def read(self, want):
        """-> (bytes, actual_len, is_end_of_stream)"""
        if not self.is_open:
            self.index += 1                                                # :48
            if not self._open(self.index):
                self.is_open = False
                return b"", 0, True                                        # :55-59
        eos = False
        if self.pos + want > self.max_len:                                 # :85
            out = self.data[self.pos:self.max_len]
            self.is_open = False
            remain = want - len(out)
            if remain > 0:
                self.index += 1                                            # :101
                if not self._open(self.index):
                    self.pos, self.max_len = 0, 0
                    return out, len(out), True                             # :106-113
                part2 = self.data[:remain]
                out += part2
                self.pos += len(part2)                                     # :133
        elif self.pos + want == self.max_len:                              # :137
            out = self.data[self.pos:self.pos + want]
            self.is_open = False
            self.index += 1                                                # :148 (then :48 increments again)
            self.pos, self.max_len = 0, 0
        else:
            out = self.data[self.pos:self.pos + want]
            self.pos += len(out)
        if len(out) < want and self.is_open:
            eos = True                                                     # :160-163
        return out, len(out), eos


def frame_data_read_ddc(stream, prtNum, point_PRT, channel_num):
    """DDC branch of FrameDataRead_xzr.m:57-198 without the DBF product: returns (sig[prt, range, channel] complex,
    servo_angle, frame_no, timer_cnt, prts_read, stream_end)."""
    sig = np.zeros((prtNum, point_PRT, channel_num), dtype=np.complex128)
    servo = np.zeros(prtNum)
    frame_no = np.zeros(prtNum, dtype=np.uint64)
    timer = np.zeros(prtNum, dtype=np.uint64)
    cur = 0
    while cur < prtNum:                                                    # :57
        hb, n, eos = stream.read(64)                                       # :62
        if eos or n < 64:
            return sig, servo, frame_no, timer, cur, True
        h = np.frombuffer(hb, dtype="<u4")                                 # :70
        ch = int(h[3]) % 256                                               # :77
        pulse_data_num = int(h[6])                                         # :79
        data_type = int(h[7]) % 256                                        # :80
        if pulse_data_num <= 0:
            return sig, servo, frame_no, timer, cur, True                  # :90-94
        _, n, eos = stream.read(128)                                       # :97
        if eos or n < 128:
            return sig, servo, frame_no, timer, cur, True
        assert data_type == 1
        size, pad = ddc_payload_size(pulse_data_num, ch)                   # :109,115-119
        pb, n, eos = stream.read(size + pad)                               # :122
        if eos or n < size + pad:
            return sig, servo, frame_no, timer, cur, True
        x = unpack_ddc_i16(np.frombuffer(pb, dtype=np.uint8), pulse_data_num, ch)     # :138,150-156
        if x.shape != (point_PRT, channel_num):                            # :171-176
            return sig, servo, frame_no, timer, cur, True
        sig[cur] = x                                                       # :179-180
        servo[cur] = int(h[4]) % 65536                                     # :78,181
        frame_no[cur] = int(h[0])                                          # :74
        timer[cur] = int(h[8]) + int(h[9]) * 2 ** 32                       # :83
        cur += 1
        _, n, eos = stream.read(64)                                        # :184
        if eos or n < 64:
            return sig, servo, frame_no, timer, cur, True
    return sig, servo, frame_no, timer, cur, False
