"""Host-side mirror of the reference's function interface (same names, argument order and error
behaviour as the M files), bound to libradar_b200.so through ctypes.

This is what the MEX gateways in ``mex/`` do, written in Python because neither MATLAB nor Octave
exists in this environment; the parity tests drive the C ABI through these functions with
MATLAB-layout (column-major, split-complex double) buffers.  No signal arithmetic happens here.

    MTD_Signal = fun_MTD_produce(echo)                     MP/fun_MTD_produce.m:3
    MTD_Signal = fun_MTD_produce(echo, params)             MTD/fun_MTD_produce.m:12
    s_PC_0 = fun_lss_pulse_compression(echo, show_PC, pulse1, pulse2, pulse3)            MP/...:3
    s_PC_0 = fun_lss_pulse_compression(echo, params, show_PC, pulse1, pulse2, pulse3, p1, p2, p3)  MTD/...:17
    signal_PC = fun_pulse_compression(s0, s_echo)          MP/fun_pulse_compression.m:1
    MTD_Signal = fun_Process_MTD(ProSignal, Len_PRT, Num_PRTperFrame)   MP/fun_Process_MTD.m:3
    MTD = fun_0v_pressing(MTD)                             MP/fun_0v_pressing.m:2 (CW/ variant: div=20)
    [F, FV] = executeCFAR(mtd, refR, saveR, T_R, methR, refV, saveV, T_V, methV, n0, rFlag)  CW/executeCFAR.m:1
    F = Function_CFAR1D_sub(data, ref, save, T, method)    CW/Function_CFAR1D_sub.m:1
    F = Function_CFAR1D_sub_fixCells(data, ref, save, T, method, rows, cols)   CW/Function_CFAR1D_sub_fixCells.m:1
"""
import os
import warnings

import numpy as np

from . import waveforms as W
from ._binding import ERR_ARG, RadarB200Error
from .context import Context

_ctx = None
_plan_key = None


def default_context():
    """Lazily created context on device ``RB200_DEVICE`` (default 0), kept alive like a MEX gateway's."""
    global _ctx
    if _ctx is None:
        _ctx = Context(int(os.environ.get("RB200_DEVICE", "0")))
    return _ctx


def shutdown():
    global _ctx, _plan_key
    if _ctx is not None:
        _ctx.close()
    _ctx = None
    _plan_key = None


def _set_plan(key, builder):
    """Reference spectra stay resident across calls with the same plan (the MEX gateway caches likewise)."""
    global _plan_key
    ctx = default_context()
    if key != _plan_key:
        ctx.set_waveform(builder())
        _plan_key = key
    return ctx


def _plot_off(*flags):
    if any(bool(np.any(f)) for f in flags):
        warnings.warn("radar_b200: show_PC/show_FFT/graph plotting is not reproduced; ignoring", RuntimeWarning)


def _key(*arrs):
    return tuple(np.asarray(a).tobytes() if isinstance(a, np.ndarray) else a for a in arrs)


def fun_pulse_compression(s0, s_echo):
    return default_context().pulse_compression(s0, s_echo)[None, :]


def fun_lss_pulse_compression(echo, *args):
    echo = np.atleast_2d(echo)
    n = echo.shape[1]
    if len(args) == 4:      # MP/ 5-arg
        show_PC, pulse1, pulse2, pulse3 = args
        _plot_off(show_PC)
        p2, p3 = np.asarray(pulse2, dtype=complex).ravel(), np.asarray(pulse3, dtype=complex).ravel()
        ctx = _set_plan(("mp", n) + _key(p2, p3), lambda: W.segments_mp(n, p2, p3))
    elif len(args) == 8:    # MTD/ 9-arg
        params, show_PC, pulse1, pulse2, pulse3, q1, q2, q3 = args
        _plot_off(show_PC)
        p2, p3 = np.asarray(pulse2, dtype=complex).ravel(), np.asarray(pulse3, dtype=complex).ravel()
        ctx = _set_plan(("mtd", n, int(q1), int(q2), int(q3)) + _key(p2, p3), lambda: W.segments_mtd(n, p2, p3, q1, q2, q3))
    else:
        raise RadarB200Error(ERR_ARG, "fun_lss_pulse_compression: expected 5 or 9 arguments")
    return ctx.lss_pulse_compression(echo)


def fun_Process_MTD(ProSignal, Len_PRT, Num_PRTperFrame):
    return default_context().process_mtd(ProSignal, Len_PRT, Num_PRTperFrame, 8.0)


def fun_0v_pressing(MTD, div=150):
    return default_context().zero_v_pressing(MTD, div)


def fun_MTD_produce(echo, params=None):
    echo = np.atleast_2d(echo)
    n = echo.shape[1]
    if params is None:      # MP/fun_MTD_produce.m: literal pulses, 5-arg PC
        ctx = _set_plan(("mp", n) + _key(W.PULSE2, W.PULSE3), lambda: W.segments_mp(n, W.PULSE2, W.PULSE3))
    else:                   # MTD/fun_MTD_produce.m: ideal LFM pulses, 9-arg PC
        dbg = params.get("debug", {})
        _plot_off(dbg.get("show_PC", 0), dbg.get("show_FFT", 0), dbg.get("graph", 0))
        _, p2, p3 = W.ideal_pulses(params)
        pp = [int(v) for v in np.ravel(params["point_prt"])]
        ctx = _set_plan(("mtd", n, pp[1], pp[2], pp[3]) + _key(p2, p3), lambda: W.segments_mtd(n, p2, p3, pp[1], pp[2], pp[3]))
    return ctx.mtd_produce(echo, 8.0, 150)


def fun_MTD_produce_rows(echo, row_lo, row_hi):
    """``MTD = fun_MTD_produce(echo); MTD = MTD(row_lo:row_hi, :)`` (MP/main_produce_dataset_win_xzr.m:37-40, rows 691:845) in
    one call that pulse-compresses only the kept Doppler rows (the slow-time transform runs first; both are linear and act
    on different axes)."""
    echo = np.atleast_2d(echo)
    n = echo.shape[1]
    ctx = _set_plan(("mp", n) + _key(W.PULSE2, W.PULSE3), lambda: W.segments_mp(n, W.PULSE2, W.PULSE3))
    return ctx.mtd_produce_rows(echo, row_lo, row_hi, 8.0, 150)


def fun_MTD_produce_windows(echo_win, win_len=1536, win_size=4):
    """The window loop of MP/main_produce_dataset_win_xzr.m:31-38 in one call: window i covers rows
    round(i*win_len/win_size)+1 ... +win_len of ``echo_win`` (two concatenated frames); pulse compression
    is done once for all rows.  Returns ``MTD[i, :, :]`` = ``fun_MTD_produce(echo_win(rows_i, :))``."""
    echo_win = np.atleast_2d(echo_win)
    n = echo_win.shape[1]
    ctx = _set_plan(("mp", n) + _key(W.PULSE2, W.PULSE3), lambda: W.segments_mp(n, W.PULSE2, W.PULSE3))
    starts = [int(np.floor(i * win_len / win_size + 0.5)) for i in range(win_size)]      # MATLAB round
    return ctx.mtd_produce_windows(echo_win, win_len, starts, 8.0, 150)


def motionParaMeasure(echo_MTD_sum_short, echo_MTD_diff_short, cfarResultFlag_Matrix_short, extraDots, rScale_short, deltaR,
                      rInterpTimes, vScale, deltaV, vInterpTimes, kValues, beamPosNum, beamAngleStep, freInd, eleAngleComp,
                      eleAngleSysErr, MTD_0_num):
    """[rEstSeries, vEstSeries, eleAngleEstSeries] = motionParaMeasure(...)   CW/motionParaMeasure.m:1"""
    return default_context().motion_para_measure(echo_MTD_sum_short, echo_MTD_diff_short, cfarResultFlag_Matrix_short, extraDots,
                                                 rScale_short, deltaR, rInterpTimes, vScale, deltaV, vInterpTimes, kValues,
                                                 beamPosNum, beamAngleStep, freInd, eleAngleComp, eleAngleSysErr, MTD_0_num)


def executeCFAR(mtd, refCells_R, saveCells_R, T_CFAR_R, CFARmethod_R, refCells_V, saveCells_V, T_CFAR_V, CFARmethod_V,
                MTD_0_num, rCFARDetect_Flag):
    return default_context().execute_cfar(mtd, refCells_R, saveCells_R, T_CFAR_R, CFARmethod_R, refCells_V, saveCells_V,
                                          T_CFAR_V, CFARmethod_V, MTD_0_num, rCFARDetect_Flag)


def Function_CFAR1D_sub(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod):
    return default_context().cfar1d_sub(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod)


def Function_CFAR1D_sub_fixCells(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod, rowCellsFix, colCellsFix):
    return default_context().cfar1d_fix(datamatrix, refCellNum, saveCellNum, T_CFAR, CFARmethod, rowCellsFix, colCellsFix)
