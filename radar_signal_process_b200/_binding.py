"""ctypes binding of libradar_b200.so (include/radar_b200.h).  No CPU fallback: if the CUDA library is
missing or no device is present, importing / creating a context raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libradar_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_INDEX, ERR_DIM_MISMATCH, ERR_NO_WAVEFORM, ERR_UNSUPPORTED, ERR_OVERFLOW = range(8)

SEG_MF, SEG_FIR, SEG_MF_CIRC = 0, 1, 2
ALIGN_LEADING_EDGE, ALIGN_DELAYED, ALIGN_GRPDELAY = 0, 1, 2
DET_V, DET_2D = 1, 2


class RadarB200Error(RuntimeError):
    def __init__(self, status, text):
        super().__init__("libradar_b200 status %d: %s" % (status, text))
        self.status = status


class MatlabIndexError(RadarB200Error):
    """The M-code would raise 'Index exceeds array bounds' for this call."""


class MatlabDimensionError(RadarB200Error):
    """The M-code would raise a dimension-mismatch error for this call."""


class DetectionOverflow(RadarB200Error):
    """More detections than rb200_config.max_det; the list is truncated."""


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("n_prt", C.c_int32), ("n_range", C.c_int32), ("n_lanes", C.c_int32),
        ("max_cpi", C.c_int32), ("max_det", C.c_int32), ("zero_v_div", C.c_int32), ("mti_lag", C.c_int32),
        ("kaiser_beta", C.c_double),
        ("cfar_ref_r", C.c_int32), ("cfar_guard_r", C.c_int32), ("cfar_method_r", C.c_int32),
        ("cfar_ref_v", C.c_int32), ("cfar_guard_v", C.c_int32), ("cfar_method_v", C.c_int32),
        ("cfar_t_r", C.c_double), ("cfar_t_v", C.c_double),
        ("cfar_n0", C.c_int32), ("cfar_range_stage", C.c_int32), ("chunk_cpi", C.c_int32), ("reserved", C.c_int32),
    ]


class Segment(C.Structure):
    _fields_ = [
        ("in_start", C.c_int32), ("in_len", C.c_int32), ("out_start", C.c_int32), ("out_len", C.c_int32),
        ("kind", C.c_int32), ("align", C.c_int32), ("n_taps", C.c_int32), ("reserved", C.c_int32),
        ("taps_re", C.POINTER(C.c_double)), ("taps_im", C.POINTER(C.c_double)), ("scale", C.c_double),
    ]


DET_DTYPE = np.dtype([("cpi", "<u4"), ("r", "<u4"), ("v", "<u2"), ("lane", "u1"), ("kind", "u1"), ("amp", "<f4")])
assert DET_DTYPE.itemsize == 16

EXPORTS = [
    "rb200_version", "rb200_create", "rb200_destroy", "rb200_last_error", "rb200_get_config", "rb200_set_cfar",
    "rb200_set_waveform", "rb200_set_stc", "rb200_pulse_compression_z", "rb200_lss_pulse_compression_z",
    "rb200_process_mtd_z", "rb200_zero_v_pressing_d", "rb200_mtd_produce_z", "rb200_mtd_produce_rows_z", "rb200_mtd_produce_c", "rb200_mtd_produce_rows_c", "rb200_cfar1d_sub_d",
    "rb200_cfar1d_fix_d", "rb200_execute_cfar_d", "rb200_unpack_ddc_i16", "rb200_chain_i16",
    "rb200_chain_enqueue", "rb200_chain_fetch", "rb200_debug_fetch_pc", "rb200_last_device_ms",
    "rb200_last_launch_count", "rb200_set_dbf", "rb200_set_cfar_segments", "rb200_set_stage_timing", "rb200_get_stage_ms", "rb200_unpack_dbf24", "rb200_chain_dbf24", "rb200_mtd_produce_windows_z", "rb200_dmx_process_z", "rb200_motion_para_measure_d", "rb200_reader_open", "rb200_reader_close",
    "rb200_reader_last_error", "rb200_reader_state", "rb200_reader_next_frame_ddc", "rb200_reader_next_frame_dbf24",
    "rb200_shared_context_acquire", "rb200_shared_context_release", "rb200_set_plan_tag", "rb200_get_plan_tag",
    "rb200_set_debug_keep_pc", "rb200_chain_dets_device",
]

_lib = None


def load():
    """Load the shared library (raises if it has not been built: run ``make`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadarB200Error(ERR_CUDA, "%s not found -- build it with `make` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    lib.rb200_version.restype = C.c_int
    lib.rb200_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Config)]
    lib.rb200_destroy.argtypes = [vp]
    lib.rb200_last_error.argtypes = [vp]
    lib.rb200_last_error.restype = C.c_char_p
    lib.rb200_get_config.argtypes = [vp, C.POINTER(Config)]
    lib.rb200_set_cfar.argtypes = [vp, C.POINTER(Config)]
    lib.rb200_set_waveform.argtypes = [vp, C.POINTER(Segment), C.c_int]
    lib.rb200_set_stc.argtypes = [vp, dp, C.c_int]
    lib.rb200_set_dbf.argtypes = [vp, dp, dp, C.c_int]
    lib.rb200_pulse_compression_z.argtypes = [vp, dp, dp, C.c_int, dp, dp, C.c_int, dp, dp]
    lib.rb200_lss_pulse_compression_z.argtypes = [vp, dp, dp, C.c_int, C.c_int, dp, dp]
    lib.rb200_process_mtd_z.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dp]
    lib.rb200_zero_v_pressing_d.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, dp]
    lib.rb200_mtd_produce_z.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_int, dp]
    lib.rb200_mtd_produce_rows_z.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp]
    lib.rb200_mtd_produce_c.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double, C.c_int, dp]
    lib.rb200_mtd_produce_rows_c.argtypes = [vp, dp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp]
    lib.rb200_cfar1d_sub_d.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp]
    lib.rb200_cfar1d_fix_d.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                       C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int, dp]
    lib.rb200_execute_cfar_d.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                         C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp]
    lib.rb200_unpack_ddc_i16.argtypes = [vp, vp, C.c_int, vp]
    lib.rb200_chain_i16.argtypes = [vp, vp, C.c_int, vp, vp, C.POINTER(C.c_int), vp]
    lib.rb200_chain_enqueue.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.rb200_chain_fetch.argtypes = [vp, vp, C.POINTER(C.c_int)]
    lib.rb200_debug_fetch_pc.argtypes = [vp, C.c_int, vp]
    lib.rb200_last_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.rb200_last_launch_count.argtypes = [vp, C.POINTER(C.c_int)]
    lib.rb200_set_stage_timing.argtypes = [vp, C.c_int]
    lib.rb200_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.rb200_unpack_dbf24.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_int)]
    lib.rb200_set_cfar_segments.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
    lib.rb200_chain_dbf24.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_int), vp]
    lib.rb200_mtd_produce_windows_z.argtypes = [vp, dp, dp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_double, C.c_int, dp]
    lib.rb200_dmx_process_z.argtypes = [vp, dp, dp, dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_int, C.c_int, dp, C.c_int, C.c_int,
                                        dp, dp, dp, dp]
    lib.rb200_motion_para_measure_d.argtypes = [vp, dp, dp, dp, C.c_int, C.c_int, C.c_int, dp, C.c_double, C.c_int, dp, C.c_double, C.c_int,
                                                dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                                dp, dp, dp, C.c_int, C.POINTER(C.c_int)]
    lib.rb200_reader_open.argtypes = [C.POINTER(vp), C.c_char_p]
    lib.rb200_reader_close.argtypes = [vp]
    lib.rb200_reader_last_error.argtypes = [vp]
    lib.rb200_reader_last_error.restype = C.c_char_p
    lib.rb200_reader_state.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
    lib.rb200_reader_next_frame_ddc.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.rb200_reader_next_frame_dbf24.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.rb200_shared_context_acquire.argtypes = [C.POINTER(vp), C.c_int]
    lib.rb200_shared_context_release.argtypes = [C.c_int]
    lib.rb200_set_plan_tag.argtypes = [vp, C.c_uint64]
    lib.rb200_get_plan_tag.argtypes = [vp]
    lib.rb200_get_plan_tag.restype = C.c_uint64
    lib.rb200_set_debug_keep_pc.argtypes = [vp, C.c_int]
    lib.rb200_chain_dets_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(vp), C.POINTER(C.c_int)]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def default_config(**kw):
    c = Config()
    c.struct_size = C.sizeof(Config)
    c.n_prt, c.n_range, c.n_lanes, c.max_cpi, c.max_det = 64, 4096, 16, 1, 65536
    c.zero_v_div, c.mti_lag, c.kaiser_beta = 150, 0, 8.0
    c.cfar_ref_r = c.cfar_ref_v = 5
    c.cfar_guard_r = c.cfar_guard_v = 7
    c.cfar_t_r = c.cfar_t_v = 5.0
    c.cfar_method_r = c.cfar_method_v = 0
    c.cfar_n0, c.cfar_range_stage, c.chunk_cpi = 0, 1, 0
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError("unknown config field %r" % k)
        setattr(c, k, v)
    return c


def raise_for(status, ctx_handle):
    if status == OK:
        return
    text = load().rb200_last_error(ctx_handle)
    text = text.decode("utf-8", "replace") if text else ""
    cls = {ERR_INDEX: MatlabIndexError, ERR_DIM_MISMATCH: MatlabDimensionError, ERR_OVERFLOW: DetectionOverflow}.get(status, RadarB200Error)
    raise cls(status, text)


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
