"""Context: one rb200_ctx (one CUDA device, one stream) behind a Python object."""
import ctypes as C

import numpy as np

from . import _binding as B


def _split(z):
    """complex/real ndarray -> (re, im|None) contiguous float64 in MATLAB (column-major) element order."""
    z = np.asarray(z)
    if np.iscomplexobj(z):
        return np.asfortranarray(z.real, dtype=np.float64), np.asfortranarray(z.imag, dtype=np.float64)
    return np.asfortranarray(z, dtype=np.float64), None


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Context:
    def __init__(self, device=0, **cfg):
        self._lib = B.load()
        self._h = C.c_void_p()
        self.cfg = B.default_config(**cfg)
        st = self._lib.rb200_create(C.byref(self._h), int(device), C.byref(self.cfg))
        if st != B.OK:
            text = self._lib.rb200_last_error(None)
            self._h = C.c_void_p()
            raise B.RadarB200Error(st, text.decode() if text else "")
        self._keep = None
        self.device = int(device)
        self.n_out_lanes = self.cfg.n_lanes      # lanes downstream of the optional beam former

    # ---- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rb200_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        B.raise_for(st, self._h)

    # ---- configuration ----------------------------------------------------------------------------
    def set_waveform(self, segments):
        arr = (B.Segment * len(segments))()
        keep = []
        for i, s in enumerate(segments):
            taps = np.asarray(s["taps"], dtype=np.complex128).ravel()
            re = np.ascontiguousarray(taps.real)
            im = np.ascontiguousarray(taps.imag)
            keep += [re, im]
            a = arr[i]
            a.in_start, a.in_len, a.out_start, a.out_len = s["in_start"], s["in_len"], s["out_start"], s["out_len"]
            a.kind, a.align, a.n_taps = s["kind"], s["align"], taps.size
            a.taps_re, a.taps_im, a.scale = _fptr(re), _fptr(im), s.get("scale", 1.0)
        self._ck(self._lib.rb200_set_waveform(self._h, arr, len(segments)))

    def set_stc(self, stc_db):
        if stc_db is None or len(stc_db) == 0:
            self._ck(self._lib.rb200_set_stc(self._h, None, 0))
            return
        a = np.ascontiguousarray(stc_db, dtype=np.float64)
        self._ck(self._lib.rb200_set_stc(self._h, _fptr(a), a.size))

    def set_dbf(self, W):
        """DBF weighting fused into the unpack: beams = sig_C * W.' (FrameDataRead_xzr.m:158).
        ``W``: (n_beams, n_lanes) complex (``DBF_coeffs_data_C``); ``None`` switches it off."""
        if W is None:
            self._ck(self._lib.rb200_set_dbf(self._h, None, None, 0))
            self.n_out_lanes = self.cfg.n_lanes
            return
        W = np.asarray(W, dtype=np.complex128)
        assert W.ndim == 2 and W.shape[1] == self.cfg.n_lanes, "W must be (n_beams, n_lanes)"
        re, im = _split(W)
        self._ck(self._lib.rb200_set_dbf(self._h, _fptr(re), _fptr(im), W.shape[0]))
        self.n_out_lanes = W.shape[0]

    def set_cfar(self, refR, saveR, T_R, methR, refV, saveV, T_V, methV, n0, rflag):
        c = self.cfg
        c.cfar_ref_r, c.cfar_guard_r, c.cfar_t_r, c.cfar_method_r = int(refR), int(saveR), float(T_R), int(methR)
        c.cfar_ref_v, c.cfar_guard_v, c.cfar_t_v, c.cfar_method_v = int(refV), int(saveV), float(T_V), int(methV)
        c.cfar_n0, c.cfar_range_stage = int(n0), int(bool(rflag))
        self._ck(self._lib.rb200_set_cfar(self._h, C.byref(c)))

    # ---- MATLAB-layout entry points ------------------------------------------------------------------
    def pulse_compression(self, s0, s_echo):
        s0 = np.asarray(s0).ravel()
        x = np.asarray(s_echo).ravel()
        sre, sim = _split(s0)
        xre, xim = _split(x)
        n = s0.size + x.size - 1
        ore, oim = np.zeros(max(n, 0)), np.zeros(max(n, 0))
        self._ck(self._lib.rb200_pulse_compression_z(self._h, _fptr(sre), _fptr(sim), s0.size, _fptr(xre), _fptr(xim), x.size,
                                                     _fptr(ore), _fptr(oim)))
        return ore + 1j * oim

    def lss_pulse_compression(self, echo):
        echo = np.atleast_2d(echo)
        P, R = echo.shape
        re, im = _split(echo)
        ore = np.zeros((P, R), order="F")
        oim = np.zeros((P, R), order="F")
        self._ck(self._lib.rb200_lss_pulse_compression_z(self._h, _fptr(re), _fptr(im), P, R, _fptr(ore), _fptr(oim)))
        return ore + 1j * oim

    def process_mtd(self, x, len_prt, num_prt, beta=8.0):
        x = np.atleast_2d(x)
        rows, cols = x.shape
        re, im = _split(x)
        out = np.zeros((int(num_prt), int(len_prt)), order="F")
        self._ck(self._lib.rb200_process_mtd_z(self._h, _fptr(re), _fptr(im), rows, cols, int(len_prt), int(num_prt), float(beta), _fptr(out)))
        return out

    def zero_v_pressing(self, mtd, div=150):
        mtd = np.asfortranarray(np.atleast_2d(mtd), dtype=np.float64)
        out = np.zeros(mtd.shape, order="F")
        self._ck(self._lib.rb200_zero_v_pressing_d(self._h, _fptr(mtd), mtd.shape[0], mtd.shape[1], int(div), _fptr(out)))
        return out

    def mtd_produce(self, echo, beta=8.0, zero_v_div=150):
        echo = np.atleast_2d(echo)
        P, R = echo.shape
        out = np.empty((P, R), order="F")
        if np.iscomplexobj(echo):       # numpy's complex128 is interleaved: one H2D copy, no host-side split
            z = np.asfortranarray(echo, dtype=np.complex128)
            self._ck(self._lib.rb200_mtd_produce_c(self._h, z.ctypes.data_as(C.POINTER(C.c_double)), P, R, float(beta), int(zero_v_div),
                                                   _fptr(out)))
            return out
        re, im = _split(echo)
        self._ck(self._lib.rb200_mtd_produce_z(self._h, _fptr(re), _fptr(im), P, R, float(beta), int(zero_v_div), _fptr(out)))
        return out

    def mtd_produce_rows(self, echo, row_lo, row_hi, beta=8.0, zero_v_div=150):
        """fun_MTD_produce(echo)(row_lo:row_hi, :) (1-based, inclusive): slow-time transform first, PC on the kept rows only."""
        echo = np.atleast_2d(echo)
        P, R = echo.shape
        out = np.zeros((max(int(row_hi) - int(row_lo) + 1, 1), R), order="F")
        if np.iscomplexobj(echo):
            z = np.asfortranarray(echo, dtype=np.complex128)
            self._ck(self._lib.rb200_mtd_produce_rows_c(self._h, z.ctypes.data_as(C.POINTER(C.c_double)), P, R, float(beta), int(zero_v_div),
                                                        int(row_lo), int(row_hi), _fptr(out)))
            return out
        re, im = _split(echo)
        self._ck(self._lib.rb200_mtd_produce_rows_z(self._h, _fptr(re), _fptr(im), P, R, float(beta), int(zero_v_div), int(row_lo), int(row_hi),
                                                    _fptr(out)))
        return out

    def mtd_produce_windows(self, echo, win_len, row_start, beta=8.0, zero_v_div=150):
        """PC once over all rows, MTD + 0-v per window (rows row_start[i] .. row_start[i]+win_len-1, 0-based)."""
        echo = np.atleast_2d(echo)
        P, R = echo.shape
        re, im = _split(echo)
        rs = np.ascontiguousarray(row_start, dtype=np.int32)
        out = np.zeros((rs.size, int(win_len), R), dtype=np.float64)
        buf = np.zeros(rs.size * int(win_len) * R)
        self._ck(self._lib.rb200_mtd_produce_windows_z(self._h, _fptr(re), _fptr(im), P, R, int(win_len), rs.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       rs.size, float(beta), int(zero_v_div), _fptr(buf)))
        for i in range(rs.size):
            out[i] = buf[i * win_len * R:(i + 1) * win_len * R].reshape((int(win_len), R), order="F")
        return out

    def set_cfar_segments(self, segments):
        """fun_CFARflag (CW/main_cfar.m:142-161): 0-based half-open (lo, hi) column ranges; [] = whole PRT."""
        lo = np.ascontiguousarray([s[0] for s in segments], dtype=np.int32)
        hi = np.ascontiguousarray([s[1] for s in segments], dtype=np.int32)
        self._ck(self._lib.rb200_set_cfar_segments(self._h, lo.ctypes.data_as(C.POINTER(C.c_int32)), hi.ctypes.data_as(C.POINTER(C.c_int32)), len(segments)))

    def dmx_process(self, left, right, n_short, fir_taps, mf_taps, fft_num, mtd_window, mtd_fft_num, n_blank):
        """One frame of the DMX script variant (CW/DMX_SignalProcessing_main_xzr.m:332-426,462-465).

        left/right: P x n_range complex beams; returns (sum_short, diff_short, sum_long, diff_long) as
        mtd_fft_num x n_short / mtd_fft_num x fft_num real matrices (the *_short ones are None when n_short = 0).
        """
        left = np.atleast_2d(left)
        right = np.atleast_2d(right)
        if left.shape != right.shape:
            raise B.MatlabDimensionError(B.ERR_DIM_MISMATCH, "dmx_process: the two beams differ in size")
        P, R = left.shape
        lre, lim = _split(left)
        rre, rim = _split(right)
        fir = np.ascontiguousarray(fir_taps, dtype=np.float64).ravel()
        mf = np.ascontiguousarray(mf_taps).ravel()
        mre = np.ascontiguousarray(mf.real, dtype=np.float64)
        mim = np.ascontiguousarray(mf.imag, dtype=np.float64) if np.iscomplexobj(mf) else None
        w = np.ascontiguousarray(mtd_window, dtype=np.float64).ravel()
        if w.size != P:
            raise B.MatlabDimensionError(B.ERR_DIM_MISMATCH, "dmx_process: mtd_window must have one entry per PRT")
        n_short, fft_num, mtd_fft_num = int(n_short), int(fft_num), int(mtd_fft_num)
        ss = np.zeros(mtd_fft_num * n_short) if n_short else None
        ds = np.zeros(mtd_fft_num * n_short) if n_short else None
        sl = np.zeros(mtd_fft_num * fft_num)
        dl = np.zeros(mtd_fft_num * fft_num)
        self._ck(self._lib.rb200_dmx_process_z(self._h, _fptr(lre), _fptr(lim), _fptr(rre), _fptr(rim), P, R, n_short, _fptr(fir), fir.size,
                                               _fptr(mre), _fptr(mim), mre.size, fft_num, _fptr(w), mtd_fft_num, int(n_blank),
                                               _fptr(ss), _fptr(ds), _fptr(sl), _fptr(dl)))
        shp = lambda a, n: None if a is None else a.reshape((mtd_fft_num, n), order="F")
        return shp(ss, n_short), shp(ds, n_short), shp(sl, fft_num), shp(dl, fft_num)

    def motion_para_measure(self, mtd_sum, mtd_diff, flags, extraDots, rScale, deltaR, rInterpTimes, vScale, deltaV, vInterpTimes,
                            kValues, beamPosNum, beamAngleStep, freInd, eleAngleComp, eleAngleSysErr, MTD_0_num):
        s = np.asfortranarray(np.atleast_2d(mtd_sum), dtype=np.float64)
        d = np.asfortranarray(np.atleast_2d(mtd_diff), dtype=np.float64)
        f = np.asfortranarray(np.atleast_2d(flags), dtype=np.float64)
        V, R = f.shape
        rs = np.ascontiguousarray(rScale, dtype=np.float64).ravel()
        vs = np.ascontiguousarray(vScale, dtype=np.float64).ravel()
        kv = np.asfortranarray(np.atleast_2d(kValues), dtype=np.float64)
        cap = int(np.count_nonzero(f))
        outs = [np.zeros(max(cap, 1)) for _ in range(3)]
        n = C.c_int(0)
        self._ck(self._lib.rb200_motion_para_measure_d(self._h, _fptr(s), _fptr(d), _fptr(f), V, R, int(extraDots), _fptr(rs), float(deltaR),
                                                       int(rInterpTimes), _fptr(vs), float(deltaV), int(vInterpTimes), _fptr(kv), kv.shape[0],
                                                       kv.shape[1], float(beamPosNum), float(beamAngleStep), int(freInd), float(eleAngleComp),
                                                       float(eleAngleSysErr), int(MTD_0_num), _fptr(outs[0]), _fptr(outs[1]), _fptr(outs[2]),
                                                       cap, C.byref(n)))
        return tuple(o[: n.value].copy() for o in outs)

    def cfar1d_sub(self, data, ref, guard, T, method):
        d = np.asfortranarray(np.atleast_2d(data), dtype=np.float64)
        out = np.zeros(d.shape, order="F")
        self._ck(self._lib.rb200_cfar1d_sub_d(self._h, _fptr(d), d.shape[0], d.shape[1], int(ref), int(guard), float(T), int(method), _fptr(out)))
        return out

    def cfar1d_fix(self, data, ref, guard, T, method, rows_fix, cols_fix):
        d = np.asfortranarray(np.atleast_2d(data), dtype=np.float64)
        out = np.zeros(d.shape, order="F")
        rf = np.ascontiguousarray(np.atleast_1d(rows_fix), dtype=np.int32)
        cf = np.ascontiguousarray(np.atleast_1d(cols_fix), dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        self._ck(self._lib.rb200_cfar1d_fix_d(self._h, _fptr(d), d.shape[0], d.shape[1], int(ref), int(guard), float(T), int(method),
                                              rf.ctypes.data_as(ip), rf.size, cf.ctypes.data_as(ip), cf.size, _fptr(out)))
        return out

    def execute_cfar(self, mtd, refR, saveR, T_R, methR, refV, saveV, T_V, methV, n0, rflag):
        d = np.asfortranarray(np.atleast_2d(mtd), dtype=np.float64)
        f = np.zeros(d.shape, order="F")
        fv = np.zeros(d.shape, order="F")
        self._ck(self._lib.rb200_execute_cfar_d(self._h, _fptr(d), d.shape[0], d.shape[1], int(refR), int(saveR), float(T_R), int(methR),
                                                int(refV), int(saveV), float(T_V), int(methV), int(n0), int(bool(rflag)), _fptr(f), _fptr(fv)))
        return f, fv

    # ---- batched wire-format entry points ---------------------------------------------------------------
    def _cells(self, n_cpi):
        c = self.cfg
        return n_cpi * c.n_prt * c.n_range * c.n_lanes

    def unpack(self, raw, n_cpi):
        c = self.cfg
        raw = np.ascontiguousarray(raw, dtype=np.int16)
        assert raw.size == self._cells(n_cpi) * 2
        out = np.zeros((n_cpi, c.n_lanes, c.n_prt, c.n_range), dtype=np.complex64)
        self._ck(self._lib.rb200_unpack_ddc_i16(self._h, raw.ctypes.data, n_cpi, out.ctypes.data))
        return out

    def unpack_dbf24(self, payload_bytes, n_prt, n_samples, n_channels):
        """DBF-type (24-bit) PRT payloads -> complex64 [column][prt][sample] (FrameDataRead_xzr.m:130-135,163)."""
        b = np.ascontiguousarray(payload_bytes, dtype=np.uint8)
        osp = 8 - (6 * n_channels) % 8
        ncol_max = (6 * n_channels + osp) // 6 + 1
        out = np.zeros((ncol_max, n_prt, n_samples), dtype=np.complex64)
        ncol = C.c_int(0)
        self._ck(self._lib.rb200_unpack_dbf24(self._h, b.ctypes.data, int(n_prt), int(n_samples), int(n_channels), out.ctypes.data, C.byref(ncol)))
        return out.reshape(-1)[: ncol.value * n_prt * n_samples].reshape(ncol.value, n_prt, n_samples).copy()

    def chain(self, raw, n_cpi, want_rdm=True, allow_overflow=False):
        """Host-buffer chain call: numpy int16 wire array in, (rdm float32 [cpi][lane][v][r] | None, dets) out."""
        c = self.cfg
        raw = np.ascontiguousarray(raw, dtype=np.int16)
        assert raw.size == self._cells(n_cpi) * 2, "raw has the wrong number of samples for the configured geometry"
        rdm = np.zeros((n_cpi, self.n_out_lanes, c.n_prt, c.n_range), dtype=np.float32) if want_rdm else None
        dets = np.zeros(c.max_det, dtype=B.DET_DTYPE)
        n = C.c_int(0)
        st = self._lib.rb200_chain_i16(self._h, raw.ctypes.data, n_cpi, rdm.ctypes.data if want_rdm else None,
                                       dets.ctypes.data, C.byref(n), None)
        if not (st == B.ERR_OVERFLOW and allow_overflow):
            self._ck(st)
        return rdm, dets[: min(n.value, c.max_det)].copy(), n.value

    def chain_dbf24(self, payload, n_ch, n_cpi, want_rdm=True, allow_overflow=False):
        """Chain on DBF-type 24-bit PRT payloads (uint8, n_cpi x n_prt x padded PRT bytes); lanes = decoded complex columns."""
        c = self.cfg
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        rdm = np.zeros((n_cpi, c.n_lanes, c.n_prt, c.n_range), dtype=np.float32) if want_rdm else None
        dets = np.zeros(c.max_det, dtype=B.DET_DTYPE)
        n = C.c_int(0)
        st = self._lib.rb200_chain_dbf24(self._h, payload.ctypes.data, int(n_ch), n_cpi, rdm.ctypes.data if want_rdm else None,
                                         dets.ctypes.data, C.byref(n), None)
        if not (st == B.ERR_OVERFLOW and allow_overflow):
            self._ck(st)
        return rdm, dets[: min(n.value, c.max_det)].copy(), n.value

    def chain_ptr(self, raw_ptr, n_cpi, rdm_ptr, dets_ptr, stream=None):
        """Raw-pointer chain call (host or device addresses as ints); returns (status, n_det)."""
        n = C.c_int(0)
        st = self._lib.rb200_chain_i16(self._h, raw_ptr, n_cpi, rdm_ptr, dets_ptr, C.byref(n), stream)
        return st, n.value

    def chain_enqueue(self, raw_dev_ptr, n_cpi, rdm_dev_ptr, stream=None):
        self._ck(self._lib.rb200_chain_enqueue(self._h, raw_dev_ptr, n_cpi, rdm_dev_ptr, stream))

    def chain_fetch(self, allow_overflow=False):
        dets = np.zeros(self.cfg.max_det, dtype=B.DET_DTYPE)
        n = C.c_int(0)
        st = self._lib.rb200_chain_fetch(self._h, dets.ctypes.data, C.byref(n))
        if not (st == B.ERR_OVERFLOW and allow_overflow):
            self._ck(st)
        return dets[: min(n.value, self.cfg.max_det)].copy(), n.value

    def chain_dets_device(self, allow_overflow=False):
        """Device pointers and counts of the last chain call's lists: ((ptr_2d, n_2d), (ptr_v, n_v)); no copy."""
        p2, pv = C.c_void_p(), C.c_void_p()
        n2, nv = C.c_int(0), C.c_int(0)
        st = self._lib.rb200_chain_dets_device(self._h, C.byref(p2), C.byref(n2), C.byref(pv), C.byref(nv))
        if not (st == B.ERR_OVERFLOW and allow_overflow):
            self._ck(st)
        return (p2.value, n2.value), (pv.value, nv.value)

    def set_debug_keep_pc(self, on=True):
        """Run later chain calls with the pulse-compressed intermediate in device memory (needed by debug_fetch_pc)."""
        self._ck(self._lib.rb200_set_debug_keep_pc(self._h, int(bool(on))))

    def debug_fetch_pc(self, cpi_in_chunk=0):
        c = self.cfg
        out = np.zeros((self.n_out_lanes, c.n_prt, c.n_range), dtype=np.complex64)
        self._ck(self._lib.rb200_debug_fetch_pc(self._h, int(cpi_in_chunk), out.ctypes.data))
        return out

    def last_device_ms(self):
        ms = C.c_float(0)
        self._ck(self._lib.rb200_last_device_ms(self._h, C.byref(ms)))
        return ms.value

    def set_stage_timing(self, on):
        self._ck(self._lib.rb200_set_stage_timing(self._h, int(bool(on))))

    def get_stage_ms(self):
        """({'pc': ms, 'mtd': ms, 'cfar': ms}, n_chunks, n_cpis) accumulated since the last call."""
        ms = (C.c_float * 3)()
        nch, ncp = C.c_int(0), C.c_int(0)
        self._ck(self._lib.rb200_get_stage_ms(self._h, ms, C.byref(nch), C.byref(ncp)))
        return {"pc": ms[0], "mtd": ms[1], "cfar": ms[2]}, nch.value, ncp.value

    def last_launch_count(self):
        n = C.c_int(0)
        self._ck(self._lib.rb200_last_launch_count(self._h, C.byref(n)))
        return n.value


def dets_to_flags(dets, n_cpi, n_lanes, V, R):
    """Rebuild executeCFAR's two dense 0/1 matrices per (cpi, lane) from a detection list."""
    flag = np.zeros((n_cpi, n_lanes, V, R), dtype=np.float64)
    flagv = np.zeros((n_cpi, n_lanes, V, R), dtype=np.float64)
    d2 = dets[(dets["kind"] & B.DET_2D) != 0]
    dv = dets[(dets["kind"] & B.DET_V) != 0]
    flag[d2["cpi"], d2["lane"], d2["v"], d2["r"]] = 1.0
    flagv[dv["cpi"], dv["lane"], dv["v"], dv["r"]] = 1.0
    return flag, flagv
