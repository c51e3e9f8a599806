"""ctypes access to the REFERENCE library built from /root/reference into
oracle/_ref/libcufinufft_ref.so (oracle/Makefile target ref_gpu) -- test infrastructure.
The built .so travels to the GPU box; /root/reference itself is never read at run time."""
import ctypes
import os
from ctypes import POINTER, byref, c_double, c_float, c_int, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libcufinufft_ref.so")


def available():
    return os.path.exists(PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH, mode=ctypes.RTLD_LOCAL)
    return _lib


class RefPlan:
    """makeplan / setpts / execute / destroy of the reference + reads of its plan struct."""

    def __init__(self, nufft_type, modes, tol, dtype, ntransf=1, maxbatch=1, iflag=None, **opts):
        from cufinufft_b200._cufinufft import NufftOpts
        self.L = lib()
        self.s = "" if np.dtype(dtype) == np.float64 else "f"
        self.dtype = np.dtype(dtype)
        real = c_double if self.s == "" else c_float
        if iflag is None:
            iflag = 1 if nufft_type == 1 else -1
        dim = len(modes)
        self.dim, self.modes = dim, tuple(modes)
        o = NufftOpts()
        fn = self.L.cufinufft_default_opts
        fn.argtypes = [c_int, c_int, POINTER(NufftOpts)]
        assert fn(nufft_type, dim, o) == 0
        for k, v in opts.items():
            setattr(o, k, v)
        m = (c_int * 3)(*(tuple(modes) + (1,) * (3 - dim)))
        self.plan = c_void_p(None)
        mk = getattr(self.L, "cufinufft%s_makeplan" % self.s)
        mk.argtypes = [c_int, c_int, POINTER(c_int), c_int, c_int, real, c_int, POINTER(c_void_p), POINTER(NufftOpts)]
        mk.restype = c_int
        ier = mk(nufft_type, dim, m, iflag, ntransf, tol, maxbatch, byref(self.plan), o)
        if ier != 0:
            raise RuntimeError("reference makeplan failed: %d" % ier)
        self.keep = []

    def _fn(self, name, argtypes):
        f = getattr(self.L, name % self.s)
        f.argtypes = argtypes
        f.restype = c_int
        return f

    def set_pts(self, dev_pts):
        """dev_pts: [x, y, z] GPUArrays (x fastest)."""
        self.keep = list(dev_pts)
        p = [a.ptr for a in dev_pts] + [None] * (3 - len(dev_pts))
        ier = self._fn("cufinufft%s_setpts", [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p])(
            dev_pts[0].size, p[0], p[1], p[2], 0, None, None, None, self.plan)
        if ier != 0:
            raise RuntimeError("reference setpts failed: %d" % ier)
        self.M = dev_pts[0].size

    def execute(self, c, fk):
        ier = self._fn("cufinufft%s_execute", [c_void_p, c_void_p, c_void_p])(c.ptr, fk.ptr, self.plan)
        if ier != 0:
            raise RuntimeError("reference execute failed: %d" % ier)

    def geometry(self):
        g = (c_int * 16)()
        assert self._fn("refg_get_ints%s", [c_void_p, c_int, c_void_p])(self.plan, 0, g) == 0
        keys = ("dim", "nf1", "nf2", "nf3", "ns", "nbins1", "nbins2", "nbins3", "binsx", "binsy", "binsz",
                "maxbatch", "M", "totalnumsubprob", "method", "nbins")
        return dict(zip(keys, list(g)))

    def bin_layout(self):
        g = self.geometry()
        out = dict(g)
        f = self._fn("refg_get_ints%s", [c_void_p, c_int, c_void_p])
        wanted = [("binsize", 1, g["nbins"]), ("binstartpts", 2, g["nbins"]), ("idxnupts", 6, g["M"])]
        if g["method"] == 2:      # the reference allocates the subproblem arrays only for the SM method
            wanted += [("numsubprob", 3, g["nbins"]), ("subprobstartpts", 4, g["nbins"] + 1),
                       ("subprob_to_bin", 5, g["totalnumsubprob"])]
        for name, what, n in wanted:
            arr = np.zeros(max(n, 1), np.int32)
            assert f(self.plan, what, arr.ctypes.data_as(c_void_p)) == 0, name
            out[name] = arr[:n]
        return out

    def phihat(self, d):
        g = self.geometry()
        arr = np.zeros(g["nf%d" % (d + 1)] // 2 + 1, self.dtype)
        assert self._fn("refg_get_reals%s", [c_void_p, c_int, c_void_p])(self.plan, d, arr.ctypes.data_as(c_void_p)) == 0
        return arr

    def spread(self, c, fw):
        assert self._fn("refg_spread%s", [c_void_p, c_void_p, c_void_p])(self.plan, c.ptr, fw.ptr) == 0

    def interp(self, c, fw):
        assert self._fn("refg_interp%s", [c_void_p, c_void_p, c_void_p])(self.plan, c.ptr, fw.ptr) == 0

    def destroy(self):
        if self.plan is not None:
            self._fn("cufinufft%s_destroy", [c_void_p])(self.plan)
            self.plan = None
