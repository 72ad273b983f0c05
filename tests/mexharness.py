"""Drive the MEX gateways of mex/ without MATLAB/Octave: build mxArrays through the shim runtime
(mex/shim/mex_shim.c), call mexFunction, convert the results back to NumPy."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "mex", "build")


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__("%s: %s" % (ident, msg))
        self.ident = ident


_shim = None


def shim():
    global _shim
    if _shim is None:
        s = C.CDLL(os.path.join(BUILD, "librbmexshim.so"), mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        s.mxCreateDoubleMatrix.restype = vp
        s.mxCreateDoubleMatrix.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        s.mxGetPr.restype = C.POINTER(C.c_double)
        s.mxGetPr.argtypes = [vp]
        s.mxGetPi.restype = C.POINTER(C.c_double)
        s.mxGetPi.argtypes = [vp]
        s.mxGetM.restype = C.c_size_t
        s.mxGetM.argtypes = [vp]
        s.mxGetN.restype = C.c_size_t
        s.mxGetN.argtypes = [vp]
        s.mxIsComplex.argtypes = [vp]
        s.mxDestroyArray.argtypes = [vp]
        s.rbshim_create_struct.restype = vp
        s.rbshim_set_field.argtypes = [vp, C.c_char_p, vp]
        s.rbshim_call.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
        s.rbshim_last_error_id.restype = C.c_char_p
        s.rbshim_last_error_msg.restype = C.c_char_p
        s.rbshim_last_warning_id.restype = C.c_char_p
        _shim = s
    return _shim


def to_mx(x):
    s = shim()
    if isinstance(x, dict):
        st = s.rbshim_create_struct()
        for k, v in x.items():
            s.rbshim_set_field(st, k.encode(), to_mx(v))
        return st
    a = np.asarray(x)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)            # MATLAB row vector
    m, n = a.shape
    cplx = np.iscomplexobj(a)
    mx = s.mxCreateDoubleMatrix(m, n, 1 if cplx else 0)
    if m * n:
        f = np.asfortranarray(a)
        re = np.ascontiguousarray(f.real.ravel(order="F"), dtype=np.float64)
        C.memmove(s.mxGetPr(mx), re.ctypes.data, re.nbytes)
        if cplx:
            im = np.ascontiguousarray(f.imag.ravel(order="F"), dtype=np.float64)
            C.memmove(s.mxGetPi(mx), im.ctypes.data, im.nbytes)
    return mx


def from_mx(mx):
    s = shim()
    m, n = s.mxGetM(mx), s.mxGetN(mx)
    out = np.zeros(m * n)
    if m * n:
        C.memmove(out.ctypes.data, s.mxGetPr(mx), out.nbytes)
    if s.mxIsComplex(mx):
        im = np.zeros(m * n)
        if m * n:
            C.memmove(im.ctypes.data, s.mxGetPi(mx), im.nbytes)
        out = out + 1j * im
    return out.reshape((m, n), order="F")


class Mex:
    """Callable wrapper of one gateway: ``Mex("executeCFAR")(mtd, 5, 7, ..., nargout=2)``."""

    def __init__(self, name):
        shim()
        self.lib = C.CDLL(os.path.join(BUILD, name + ".so"))
        self.fn = C.cast(self.lib.mexFunction, C.c_void_p)
        self.name = name

    def __call__(self, *args, nargout=1):
        s = shim()
        prhs = (C.c_void_p * max(len(args), 1))(*[to_mx(a) for a in args])
        plhs = (C.c_void_p * max(nargout, 1))()
        rc = s.rbshim_call(self.fn, nargout, plhs, len(args), prhs)
        for i in range(len(args)):
            s.mxDestroyArray(prhs[i])
        if rc:
            raise MexError(s.rbshim_last_error_id().decode(), s.rbshim_last_error_msg().decode())
        outs = [from_mx(plhs[i]) for i in range(max(nargout, 1))]
        for i in range(max(nargout, 1)):
            s.mxDestroyArray(plhs[i])
        return outs[0] if nargout <= 1 else tuple(outs)
