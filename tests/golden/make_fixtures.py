#!/usr/bin/env python3
"""Regenerate the committed known-answer fixtures from the read-only reference tree.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_fixtures.py

What it extracts (data only, no code):

* ``kaiser_win_1536.npy``   <- MatlabProcess_xuzerui/kaiser_win.mat  (``kaiser(1536,8)``; the only
  MATLAB-computed *output* the reference ships -- SURVEY.md section 4)
* ``refDDCDataMF1.npy``     <- MatlabProcess_xuzerui/refDDCDataMF1.mat (67 complex taps)
* ``refDBFDataMF1.npy``     <- MatlabProcess_xuzerui/refDBFDataMF1.mat (67 complex taps)
* ``pulse_literals.npz``    <- literal vectors at MatlabProcess_xuzerui/fun_MTD_produce.m:54-59 and
  MatlabProcess_xuzerui/fun_lss_pulse_compression.m:21
* ``sha256.json``           digest of every fixture, checked by tests/test_oracle_golden.py
"""
import hashlib
import json
import os
import re
import sys

import numpy as np
import scipy.io as sio

REF = "/root/reference"
MP = os.path.join(REF, "MatlabProcess_xuzerui")
HERE = os.path.dirname(os.path.abspath(__file__))


def _read_m(path):
    raw = open(path, "rb").read()
    for enc in ("utf-8", "gb18030"):
        try:
            return raw.decode(enc)
        except UnicodeDecodeError:
            continue
    return raw.decode("gb18030", errors="ignore")


def _vector(src, name):
    m = re.search(r"^\s*%s\s*=\s*\[([^\]]*)\]" % re.escape(name), src, re.M)
    if not m:
        raise SystemExit("literal %s not found" % name)
    return np.array([float(t) for t in m.group(1).replace(",", " ").split()], dtype=np.float64)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are committed, nothing to do")
    out = {}
    k = sio.loadmat(os.path.join(MP, "kaiser_win.mat"))["kaiser_win"].ravel().astype(np.float64)
    out["kaiser_win_1536.npy"] = k
    out["refDDCDataMF1.npy"] = sio.loadmat(os.path.join(MP, "refDDCDataMF1.mat"))["refData"].ravel().astype(np.complex128)
    out["refDBFDataMF1.npy"] = sio.loadmat(os.path.join(MP, "refDBFDataMF1.mat"))["refData"].ravel().astype(np.complex128)
    for fn, arr in out.items():
        np.save(os.path.join(HERE, fn), arr)

    src = _read_m(os.path.join(MP, "fun_MTD_produce.m"))
    p2 = _vector(src, "pulse2_real") + 1j * _vector(src, "pulse2_imag")
    p3 = _vector(src, "pulse3_real") + 1j * _vector(src, "pulse3_imag")
    fir = _vector(_read_m(os.path.join(MP, "fun_lss_pulse_compression.m")), "filter_coef")
    assert p2.size == 75 and p3.size == 160 and fir.size == 35, (p2.size, p3.size, fir.size)
    np.savez(os.path.join(HERE, "pulse_literals.npz"), pulse2=p2, pulse3=p3, filter_coef=fir)

    digests = {}
    for fn in sorted(list(out) + ["pulse_literals.npz"]):
        if fn.endswith(".npz"):
            z = np.load(os.path.join(HERE, fn))
            h = hashlib.sha256()
            for key in sorted(z.files):
                h.update(key.encode())
                h.update(np.ascontiguousarray(z[key]).tobytes())
            digests[fn] = h.hexdigest()
        else:
            digests[fn] = hashlib.sha256(np.ascontiguousarray(np.load(os.path.join(HERE, fn))).tobytes()).hexdigest()
    json.dump(digests, open(os.path.join(HERE, "sha256.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(digests, indent=1))


if __name__ == "__main__":
    main()
