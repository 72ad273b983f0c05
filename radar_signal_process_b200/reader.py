"""Capture-file reader: ``<dir>/1.00000N.bin`` -> int16 DDC frames in the wire
layout of the chain.  Thin ctypes wrapper over rb200_reader_* (host C++ in libradar_b200.so)."""
import ctypes as C

import numpy as np

from . import _binding as B


class FrameReader:
    def __init__(self, directory):
        self._lib = B.load()
        self._h = C.c_void_p()
        st = self._lib.rb200_reader_open(C.byref(self._h), str(directory).encode())
        if st != B.OK:
            raise B.RadarB200Error(st, "cannot open capture directory %r" % (directory,))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rb200_reader_close(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def state(self):
        idx, pos = C.c_int(0), C.c_longlong(0)
        self._lib.rb200_reader_state(self._h, C.byref(idx), C.byref(pos))
        return idx.value, pos.value

    def next_frame(self, n_prt, n_range, n_channels, out=None):
        """Returns (raw[prt][range][channel][2] int16, meta dict, prts_read, end_of_stream)."""
        if out is None:
            out = np.zeros((n_prt, n_range, n_channels, 2), dtype=np.int16)
        assert out.dtype == np.int16 and out.size == n_prt * n_range * n_channels * 2 and out.flags["C_CONTIGUOUS"]
        frame_no = np.zeros(n_prt, dtype=np.uint32)
        servo = np.zeros(n_prt, dtype=np.uint16)
        timer = np.zeros(n_prt, dtype=np.uint64)
        nread, eos = C.c_int(0), C.c_int(0)
        st = self._lib.rb200_reader_next_frame_ddc(self._h, n_prt, n_range, n_channels, out.ctypes.data, frame_no.ctypes.data,
                                                   servo.ctypes.data, timer.ctypes.data, C.byref(nread), C.byref(eos))
        if st != B.OK:
            msg = self._lib.rb200_reader_last_error(self._h)
            raise B.RadarB200Error(st, msg.decode() if msg else "")
        return out, dict(frame_no=frame_no, servo_angle=servo, timer_cnt=timer), nread.value, bool(eos.value)

    def next_frame_dbf24(self, n_prt, n_range, n_channels, out=None):
        """DBF-type frames (data_type 2): returns (payload[prt][padded PRT bytes] uint8, meta, prts_read, end_of_stream);
        the payload is what Context.chain_dbf24 / unpack_dbf24 take."""
        w = 6 * n_channels + (8 - (6 * n_channels) % 8)
        sig = n_range * w
        prt_bytes = sig + ((64 - sig % 64) % 64)
        if out is None:
            out = np.zeros((n_prt, prt_bytes), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.size == n_prt * prt_bytes and out.flags["C_CONTIGUOUS"]
        frame_no = np.zeros(n_prt, dtype=np.uint32)
        servo = np.zeros(n_prt, dtype=np.uint16)
        timer = np.zeros(n_prt, dtype=np.uint64)
        nread, eos = C.c_int(0), C.c_int(0)
        st = self._lib.rb200_reader_next_frame_dbf24(self._h, n_prt, n_range, n_channels, out.ctypes.data, frame_no.ctypes.data,
                                                     servo.ctypes.data, timer.ctypes.data, C.byref(nread), C.byref(eos))
        if st != B.OK:
            msg = self._lib.rb200_reader_last_error(self._h)
            raise B.RadarB200Error(st, msg.decode() if msg else "")
        return out, dict(frame_no=frame_no, servo_angle=servo, timer_cnt=timer), nread.value, bool(eos.value)
