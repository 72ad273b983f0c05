"""CPU oracle for the PC -> MTD -> 0-v -> CFAR hot path.  TEST INFRASTRUCTURE ONLY.

This package is a double-precision NumPy restatement of the reference M-code
(XuZerui2023/Radar-Signal-Process).  It exists so that the CUDA path can be checked against the
reference's algorithm.  It is **not** part of the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import it;
* nothing under ``radar_signal_process_b200/`` imports it, and the product raises if the CUDA
  library is missing -- there is no CPU fallback.

PARITY PINNING STATUS: **parity unpinned** for the chain outputs.  The reference ships no tests and
no golden outputs, and neither MATLAB nor GNU Octave exists in the build
container or on the GPU boxes, so the M-code itself cannot be run.  What *is* pinned:

* ``kaiser(1536, 8)`` against ``MatlabProcess_xuzerui/kaiser_win.mat`` (MATLAB-computed),
* the literal matched-filter pulses / FIR taps / captured reference chirps (inputs, bit-identical),
* analytic known answers (impulse -> conj-reversed reference, single tone -> one Doppler bin,
  constant background -> CFAR rule, isolated spike -> one 2-D detection),
* loop-faithful transcription (``oracle.mcode``) == vectorised twin (``oracle.vec``).

Modules
-------
``mcode``  line-by-line transcription of the M functions (same loops, 1-based index arithmetic).
``vec``    vectorised twin used for large shapes and as the timed CPU baseline.
``synth``  seeded synthetic inputs S1..S5 of SURVEY.md section 8(d) and the int16 wire writer.
"""
from . import mcode, vec, synth  # noqa: F401
