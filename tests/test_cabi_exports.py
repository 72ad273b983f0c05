"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares;
host-side plan logic behaves like the M-code's argument handling.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import radar_signal_process_b200 as rsp
from oracle import mcode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "radar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(rsp.LIB_PATH):
        subprocess.check_call(["make", "-j8"], cwd=ROOT)
    lib = ctypes.CDLL(rsp.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(rsp.EXPORTS) == names
    assert lib.rb200_version() == 1


def test_config_struct_layout_matches_header():
    # sizeof(rb200_config) is checked by rb200_create(struct_size); here the Python mirror's size is pinned
    assert ctypes.sizeof(rsp._binding.Config) == 96
    assert ctypes.sizeof(rsp._binding.Segment) == 56
    assert rsp.DET_DTYPE.itemsize == 16


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rsp.RadarB200Error) as e:
        rsp.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "radar_signal_process_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_literal_constants_match_reference_fixtures():
    p2, p3 = mcode.load_pulse_literals()
    assert np.array_equal(rsp.waveforms.PULSE2, p2) and np.array_equal(rsp.waveforms.PULSE3, p3)
    assert np.array_equal(rsp.waveforms.FILTER_COEF, mcode.FILTER_COEF_INT)
    assert np.array_equal(rsp.waveforms.pulse1_mp(), mcode.pulse1_mp())
    params = dict(fs=25e6, B=20e6, tao=[0.16e-6, 8e-6, 28e-6], point_prt=[3404, 228, 723, 2453])
    for a, b in zip(rsp.waveforms.ideal_pulses(params), mcode.ideal_pulses_mtd(params)):
        assert np.array_equal(a, b)


def test_plan_builders_follow_the_mcode_rules():
    W = rsp.waveforms
    s = W.segments_mp(1031, W.PULSE2, W.PULSE3)
    assert [(x["in_start"], x["in_len"]) for x in s] == [(0, 82), (82, 242), (324, 707)]
    assert s[0]["kind"] == rsp._binding.SEG_FIR and s[0]["align"] == rsp._binding.ALIGN_DELAYED and abs(s[0]["scale"] - 1 / 1.2) < 1e-16
    with pytest.raises(rsp.MatlabDimensionError):
        W.segments_mp(1031, W.PULSE2[:74], W.PULSE3)
    with pytest.raises(rsp.MatlabIndexError):
        W.segments_mp(300, W.PULSE2, W.PULSE3)
    s = W.segments_mtd(3404, np.ones(200), np.ones(700), 228, 723, 2000)
    assert (s[2]["in_len"], s[2]["out_len"]) == (2453, 2000) and s[0]["align"] == rsp._binding.ALIGN_GRPDELAY
    with pytest.raises(rsp.MatlabIndexError):
        W.segments_mtd(3404, np.ones(200), np.ones(700), 228, 723, 2454)


def test_dets_to_flags_roundtrip():
    d = np.zeros(3, dtype=rsp.DET_DTYPE)
    d["cpi"], d["lane"], d["v"], d["r"], d["kind"] = [0, 1, 1], [0, 2, 2], [5, 6, 6], [7, 8, 8], [1, 2, 1]
    f, fv = rsp.dets_to_flags(d, 2, 3, 10, 10)
    assert f.sum() == 1 and f[1, 2, 6, 8] == 1 and fv.sum() == 2 and fv[0, 0, 5, 7] == 1


def test_host_radix_butterflies():
    exe = "/tmp/rb200_test_radix"
    subprocess.check_call(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tests", "host", "test_radix.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
