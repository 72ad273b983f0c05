"""radar_signal_process_b200 -- B200-native (sm_100a) PC -> MTD -> 0-v -> CFAR detection chain.

The product is ``libradar_b200.so`` (hand-written CUDA behind the C ABI of ``include/radar_b200.h``).
This package is the Python host mirror of the reference's MATLAB function interface plus the batched
wire-format chain used by the benchmark and the multi-GPU path.  It never computes on the CPU: if
the CUDA library is missing or no device is present, calls raise.
"""
from ._binding import (DET_2D, DET_DTYPE, DET_V, DetectionOverflow, MatlabDimensionError, MatlabIndexError,  # noqa: F401
                       RadarB200Error, LIB_PATH, EXPORTS, load, default_config)
from .context import Context, dets_to_flags  # noqa: F401
from . import waveforms  # noqa: F401
from .matlab_api import (fun_MTD_produce, fun_lss_pulse_compression, fun_pulse_compression, fun_Process_MTD,  # noqa: F401
                         fun_0v_pressing, executeCFAR, Function_CFAR1D_sub, Function_CFAR1D_sub_fixCells,
                         fun_MTD_produce_windows, fun_MTD_produce_rows, motionParaMeasure, default_context, shutdown)

__all__ = ["Context", "fun_MTD_produce", "fun_lss_pulse_compression", "fun_pulse_compression", "fun_Process_MTD",
           "fun_0v_pressing", "executeCFAR", "Function_CFAR1D_sub", "Function_CFAR1D_sub_fixCells", "dets_to_flags",
           "waveforms", "RadarB200Error", "MatlabIndexError", "MatlabDimensionError", "DetectionOverflow"]
