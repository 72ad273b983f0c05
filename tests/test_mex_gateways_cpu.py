"""The MEX gateways build against the shim, export mexFunction, and reject wrong argument counts the
way the interpreter would -- all without touching the GPU (those checks precede context creation)."""
import os
import subprocess

import numpy as np
import pytest

from mexharness import BUILD, Mex, MexError, ROOT

NAMES = ["motionParaMeasure", "fun_MTD_produce", "fun_MTD_produce_rows", "fun_lss_pulse_compression", "fun_pulse_compression", "fun_Process_MTD", "fun_0v_pressing",
         "fun_0v_pressing_cw", "executeCFAR", "Function_CFAR1D_sub", "Function_CFAR1D_sub_fixCells"]


@pytest.fixture(scope="module", autouse=True)
def built():
    if not all(os.path.exists(os.path.join(BUILD, n + ".so")) for n in NAMES):
        subprocess.check_call(["make", "-j8", "all"], cwd=ROOT)


@pytest.mark.parametrize("name", NAMES)
def test_gateway_loads_and_exports_mexfunction(name):
    assert Mex(name).fn


@pytest.mark.parametrize("name,nargs,ident", [
    ("fun_MTD_produce", 0, "radar_b200:mtdproduce:nargin"),
    ("fun_MTD_produce", 3, "radar_b200:mtdproduce:nargin"),
    ("fun_MTD_produce_rows", 1, "radar_b200:mtdproduce:nargin"),
    ("fun_lss_pulse_compression", 4, "radar_b200:pc:nargin"),
    ("fun_pulse_compression", 1, "radar_b200:pc:nargin"),
    ("fun_Process_MTD", 2, "radar_b200:mtd:nargin"),
    ("fun_0v_pressing", 2, "radar_b200:zerov:nargin"),
    ("executeCFAR", 10, "radar_b200:cfar:nargin"),
    ("Function_CFAR1D_sub", 4, "radar_b200:cfar1d:nargin"),
    ("Function_CFAR1D_sub_fixCells", 5, "radar_b200:cfar1d:nargin"),
    ("motionParaMeasure", 16, "radar_b200:measure:nargin"),
])
def test_wrong_nargin_raises_like_matlab(name, nargs, ident):
    with pytest.raises(MexError) as e:
        Mex(name)(*[np.ones((2, 2))] * nargs)
    assert e.value.ident == ident


def test_too_many_outputs_rejected():
    with pytest.raises(MexError) as e:
        Mex("executeCFAR")(*[np.ones((2, 2))] * 11, nargout=3)
    assert e.value.ident == "radar_b200:cfar:nargout"


def test_gateway_sources_cite_the_reference_and_bind_only_declared_symbols():
    import re
    hdr = open(os.path.join(ROOT, "include", "radar_b200.h")).read()
    declared = set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    for fn in os.listdir(os.path.join(ROOT, "mex")):
        if not fn.endswith((".cpp", ".h")) or fn == "rb200_waveform_literals.h":
            continue
        src = open(os.path.join(ROOT, "mex", fn)).read()
        used = set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", src))
        assert used <= declared, (fn, used - declared)
        if fn.endswith(".cpp"):
            assert ".m:" in src, "%s must cite the reference file:line it replaces" % fn
