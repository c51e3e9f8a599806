"""High-level plan class, a drop-in for the reference's python/cufinufft/cufinufft.py:34-267.

Same constructor signature, same `set_pts` / `execute` methods, same exceptions and
messages (python/cufinufft/tests/test_error_checks.py passes unchanged against it).
Arrays are anything exposing `.ptr` (device address), `.dtype` and `.size`:
pycuda GPUArrays as in the reference, or cufinufft_b200.gpuarray.GPUArray.
Extras beyond the reference: `maxbatch` argument, stage-level `spread`/`interp`,
`bin_layout()`, `phihat()`, `timing()`.
"""
import atexit
import ctypes
from ctypes import byref, c_int, c_void_p

import numpy as np

from . import _cufinufft as _ll

_interpreter_alive = [True]
atexit.register(_interpreter_alive.__setitem__, 0, False)


class cufinufft:
    def __init__(self, nufft_type, modes, n_trans=1, eps=1e-6, isign=None, dtype=np.float32, maxbatch=1, **kwargs):
        """
        :param nufft_type: 1 or 2 (3 is not implemented, as in the reference).
        :param modes: shape of the uniform array, C order: (nZ, nY, nX) / (nY, nX) / (nX,).
        :param n_trans: number of transforms sharing the points.
        :param eps: requested tolerance.
        :param isign: sign of the exponent; default +1 for type 1, -1 for type 2.
        :param dtype: np.float32 or np.float64.
        :param maxbatch: transforms per internal batch (the reference binding hard-wires 1;
                         0 = library heuristic min(n_trans, 8)).
        :param kwargs: fields of the options struct (gpu_method, gpu_sort, gpu_kerevalmeth, ...).
        """
        self.plan = None            # set first: __del__ must work after a failed constructor
        self.references = []
        if isign is None:
            isign = -1 if nufft_type == 2 else +1
        self.dtype = np.dtype(dtype)
        if self.dtype == np.float64:
            self._fn, self.complex_dtype = _ll._api[""], np.complex128
        elif self.dtype == np.float32:
            self._fn, self.complex_dtype = _ll._api["f"], np.complex64
        else:
            raise TypeError("Expected np.float32 or np.float64.")
        self.dim = len(modes)
        self._finufft_type = nufft_type
        self.isign = isign
        self.eps = float(eps)
        self.n_trans = n_trans
        self._maxbatch = maxbatch
        # (nZ, nY, nX) -> (nX, nY, nZ), padded with ones: the library always reads 3 ints
        self.modes = (c_int * 3)(*(tuple(modes)[::-1] + (1,) * (3 - self.dim)))
        self.opts = self._default_opts(nufft_type, self.dim)
        known = {name for name, _ in self.opts._fields_}
        for key, value in kwargs.items():
            if key not in known:
                raise TypeError(f"Invalid option '{key}'")
            setattr(self.opts, key, value)
        self._plan()

    @staticmethod
    def _default_opts(nufft_type, dim):
        opts = _ll.NufftOpts()
        if _ll._default_opts(nufft_type, dim, opts) != 0:
            raise RuntimeError('Configuration not yet implemented.')
        return opts

    def _plan(self):
        handle = c_void_p(None)
        ier = self._fn["make_plan"](self._finufft_type, self.dim, self.modes, self.isign, self.n_trans, self.eps,
                                    self._maxbatch, byref(handle), self.opts)
        if ier != 0:
            raise RuntimeError('Error creating plan.')
        self.plan = handle

    def set_pts(self, kx, ky=None, kz=None):
        """Set the nonuniform points (device arrays of the plan's real dtype, kept referenced)."""
        for name, arr in (("kx", kx), ("ky", ky), ("kz", kz)):
            if arr is not None and arr.dtype != self.dtype:
                raise TypeError("cufinufft plan.dtype and %s dtypes do not match." % name)
        M = kx.size
        if ky is not None and ky.size != M:
            raise TypeError("Number of elements in kx and ky must be equal")
        if kz is not None and kz.size != M:
            raise TypeError("Number of elements in kx and kz must be equal")
        # Python arrays are C-ordered, the library is x-fastest: the LAST python axis is the
        # library's x.  (x) -> (x); (x,y) -> (y,x); (x,y,z) -> (z,y,x)  [reference :208-219]
        given = [a for a in (kx, ky, kz) if a is not None]
        self.references = list(given)
        ptrs = [a.ptr for a in reversed(given)] + [None] * (3 - len(given))
        ier = self._fn["set_pts"](M, ptrs[0], ptrs[1], ptrs[2], 0, None, None, None, self.plan)
        if ier != 0:
            raise RuntimeError('Error setting non-uniform points.')
        self.M = M

    def execute(self, c, fk):
        """Type 1: c -> fk.  Type 2: fk -> c.  Device arrays of the plan's complex dtype."""
        if not c.dtype == fk.dtype == self.complex_dtype:
            raise TypeError("cufinufft execute expects {} dtype arguments "
                            "for this plan. Check plan and arguments.".format(self.complex_dtype))
        if self._fn["exec_plan"](c.ptr, fk.ptr, self.plan) != 0:
            raise RuntimeError('Error executing plan.')

    # ---- extras ---------------------------------------------------------------
    def set_stream(self, stream_handle):
        if self._fn["set_stream"](self.plan, c_void_p(int(stream_handle))) != 0:
            raise RuntimeError('Error setting stream.')

    def set_pts_host(self, kx, ky=None, kz=None):
        """set_pts with HOST numpy arrays (cufinufft[f]_setpts_host): copied to a plan-owned device buffer."""
        given = [np.ascontiguousarray(a, self.dtype) for a in (kx, ky, kz) if a is not None]
        M = given[0].size
        if any(a.size != M for a in given):
            raise TypeError("Number of elements in kx, ky, kz must be equal")
        self.references = list(given)
        ptrs = [a.ctypes.data for a in reversed(given)] + [None] * (3 - len(given))
        if self._fn["set_pts_host"](M, ptrs[0], ptrs[1], ptrs[2], self.plan) != 0:
            raise RuntimeError('Error setting non-uniform points.')
        self.M = M

    def execute_host(self, c, fk):
        """execute with HOST numpy arrays (cufinufft[f]_execute_host); returns when the result is in host memory."""
        if not (c.dtype == fk.dtype == self.complex_dtype and c.flags.c_contiguous and fk.flags.c_contiguous):
            raise TypeError("cufinufft execute_host expects contiguous {} arrays.".format(self.complex_dtype))
        if self._fn["exec_host"](c.ctypes.data, fk.ctypes.data, self.plan) != 0:
            raise RuntimeError('Error executing plan.')

    def spread(self, c, fw, n_trans=1):
        if self._fn["spread"](c.ptr, fw.ptr, n_trans, self.plan) != 0:
            raise RuntimeError('Error spreading.')

    def interp(self, c, fw, n_trans=1):
        if self._fn["interp"](c.ptr, fw.ptr, n_trans, self.plan) != 0:
            raise RuntimeError('Error interpolating.')

    def geometry(self):
        g = (c_int * 16)()
        if self._fn["get_ints"](self.plan, 0, g) != 0:
            raise RuntimeError('Error reading plan geometry.')
        keys = ("dim", "nf1", "nf2", "nf3", "ns", "nbins1", "nbins2", "nbins3", "binsx", "binsy", "binsz",
                "maxbatch", "M", "totalnumsubprob", "method", "nbins")
        return dict(zip(keys, list(g)))

    def bin_layout(self):
        """binsize / binstartpts / numsubprob / subprobstartpts / subprob_to_bin / idxnupts as numpy."""
        g = self.geometry()
        sizes = dict(binsize=(1, g["nbins"]), binstartpts=(2, g["nbins"]), numsubprob=(3, g["nbins"]),
                     subprobstartpts=(4, g["nbins"] + 1), subprob_to_bin=(5, g["totalnumsubprob"]),
                     idxnupts=(6, max(g["M"], 0)))
        out = dict(g)
        for name, (what, n) in sizes.items():
            arr = np.zeros(max(n, 1), np.int32)
            if self._fn["get_ints"](self.plan, what, arr.ctypes.data_as(c_void_p)) != 0:
                raise RuntimeError('Error reading %s.' % name)
            out[name] = arr[:n]
        return out

    def phihat(self, d):
        g = self.geometry()
        arr = np.zeros(g["nf%d" % (d + 1)] // 2 + 1, self.dtype)
        if self._fn["get_reals"](self.plan, d, arr.ctypes.data_as(c_void_p)) != 0:
            raise RuntimeError('Error reading phihat.')
        return arr

    def kernel_params(self):
        arr = np.zeros(3, self.dtype)
        self._fn["get_reals"](self.plan, -1, arr.ctypes.data_as(c_void_p))
        return dict(beta=arr[0], c=arr[1], halfwidth=arr[2])

    def set_timing(self, on=True):
        self._fn["set_timing"](self.plan, int(bool(on)))

    def timing(self):
        t = (ctypes.c_float * 5)()
        if self._fn["get_timing"](self.plan, t) != 0:
            raise RuntimeError('No timing recorded.')
        return dict(zip(("spread_interp_ms", "fft_ms", "deconv_amplify_ms", "memset_ms", "total_ms"), list(t)))

    def set_interp_engine(self, engine):
        """0 automatic, 1 gather engine, 2 shared-memory tile engine (include/cufinufft_b200.h)."""
        if self._fn["set_interp_engine"](self.plan, int(engine)) != 0:
            raise RuntimeError('Error selecting the interpolation engine.')

    def set_sort_levels(self, levels):
        """0 automatic, 1 one-level (bin, cell) histogram, 2 bins globally + cells per work item."""
        if self._fn["set_sort_levels"](self.plan, int(levels)) != 0:
            raise RuntimeError('Error selecting the sort mode.')

    def launch_counts(self):
        n = (c_int * 2)()
        self._fn["get_launch_counts"](self.plan, n)
        return dict(setpts=n[0], execute=n[1])

    def destroy(self):
        if self.plan is not None:
            ier = self._fn["destroy_plan"](self.plan)
            self.plan = None
            self.references = []
            if ier != 0:
                raise RuntimeError('Error destroying plan.')

    def __del__(self):
        if _interpreter_alive[0] and getattr(self, "plan", None) is not None:
            self.destroy()
