"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU oracle (oracle/liboracle.so, built from nufft_oracle.c)
plus numpy glue for the full type-1/type-2 pipeline (the FFT stage is numpy's
pocketfft, standing in for cuFFT exactly as the reference treats cuFFT as a
black box: src/2d/cufinufft2d.cu:73,141).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product never does.
"""
import ctypes
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_long, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)


def _load():
    if not os.path.exists(_LIB):
        build()
    return ctypes.CDLL(_LIB)


lib = _load()
lib.orc_next235beven.restype = c_int
lib.orc_set_nf.restype = c_int
lib.orc_set_nf.argtypes = [c_int, c_double, c_int, c_int, c_int]


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "", c_double
    if dtype == np.float32:
        return "f", c_float
    raise TypeError("dtype must be float32 or float64")


class KernelParams:
    """ns / beta / c / halfwidth for a given tol: contrib/spreadinterp.cpp:6-67."""

    def __init__(self, tol, dtype, upsampfac=2.0, kerevalmeth=0):
        s, FT = _sfx(dtype)
        ns = c_int()
        beta, hw, cc = FT(), FT(), FT()
        fn = getattr(lib, "orc_setup_spreader" + s)
        fn.argtypes = [FT, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        fn.restype = c_int
        self.ier = fn(FT(tol), upsampfac, kerevalmeth, ctypes.byref(ns), ctypes.byref(beta),
                      ctypes.byref(hw), ctypes.byref(cc))
        self.ns, self.beta, self.halfwidth, self.c = ns.value, beta.value, hw.value, cc.value
        self.dtype = np.dtype(dtype)
        self.kerevalmeth = kerevalmeth


def next235beven(n, b=1):
    return lib.orc_next235beven(int(n), int(b))


def set_nf(ms, ns, upsampfac=2.0, gpu_method=2, obinsize=1):
    return lib.orc_set_nf(int(ms), float(upsampfac), int(ns), int(gpu_method), int(obinsize))


def default_binsize(dim, gpu_method, bs=(-1, -1, -1), obs=(-1, -1, -1)):
    b = (c_int * 3)(*bs)
    o = (c_int * 3)(*obs)
    lib.orc_default_binsize(c_int(dim), c_int(gpu_method), b, o)
    return list(b), list(o)


def gauss_legendre(n):
    x = np.zeros(n)
    w = np.zeros(n)
    lib.orc_gauss_legendre(c_int(n), _p(x), _p(w))
    return x, w


def fseries_precomp(nf, kp):
    s, FT = _sfx(kp.dtype)
    q = int(2 + 3.0 * (kp.ns / 2.0))
    f = np.zeros(q, kp.dtype)
    a = np.zeros(2 * q, np.float64)
    fn = getattr(lib, "orc_fseries_precomp" + s)
    fn.argtypes = [c_int, c_int, FT, FT, FT, c_void_p, c_void_p]
    fn(nf, kp.ns, FT(kp.beta), FT(kp.c), FT(kp.halfwidth), _p(f), _p(a))
    return f, a


def fwkerhalf(nf, kp, cpu_variant=False):
    """phihat on k=0..nf/2: device formula (src/common.cu:16-45) or the reference's CPU loop."""
    s, _ = _sfx(kp.dtype)
    f, a = fseries_precomp(nf, kp)
    out = np.zeros(nf // 2 + 1, kp.dtype)
    fn = getattr(lib, ("orc_fseries_cpu" if cpu_variant else "orc_fseries_compute") + s)
    fn.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p]
    fn(nf, kp.ns, _p(f), _p(a), _p(out))
    return out


def binsort(pts, nf, bs, maxsubprobsize=1024):
    """pts: list of dim coordinate arrays; nf, bs: per-dim tuples.  Returns dict of int32 arrays."""
    dim = len(pts)
    dtype = pts[0].dtype
    s, _ = _sfx(dtype)
    M = pts[0].size
    nf = list(nf) + [1] * (3 - dim)
    bs = list(bs) + [1] * (3 - dim)
    nb = [int(np.ceil(np.asarray(nf[d], dtype) / np.asarray(bs[d], dtype))) if d < dim else 1 for d in range(3)]
    nbins = nb[0] * nb[1] * nb[2]
    out = dict(binsize=np.zeros(nbins, np.int32), binstartpts=np.zeros(nbins, np.int32),
               idxnupts=np.zeros(max(M, 1), np.int32), numsubprob=np.zeros(nbins, np.int32),
               subprobstartpts=np.zeros(nbins + 1, np.int32),
               subprob_to_bin=np.zeros(nbins + M // max(maxsubprobsize, 1) + 1, np.int32))
    fn = getattr(lib, "orc_binsort" + s)
    fn.restype = c_int
    fn.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    xs = [np.ascontiguousarray(p) for p in pts] + [None] * (3 - dim)
    T = fn(dim, M, _p(xs[0]), _p(xs[1]), _p(xs[2]), nf[0], nf[1], nf[2], bs[0], bs[1], bs[2], maxsubprobsize,
           _p(out["binsize"]), _p(out["binstartpts"]), _p(out["idxnupts"]), _p(out["numsubprob"]),
           _p(out["subprobstartpts"]), _p(out["subprob_to_bin"]))
    out["totalnumsubprob"] = T
    out["subprob_to_bin"] = out["subprob_to_bin"][:T]
    out["idxnupts"] = out["idxnupts"][:M]
    out["nbins"] = nb
    return out


def _cplx(a, dtype):
    cd = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    return np.ascontiguousarray(a, dtype=cd)


def spread(pts, c, nf, kp, fw=None):
    """fw[nf3][nf2][nf1] += sum_j c_j phi(...)  (Spread_*_NUptsdriven)."""
    dim = len(pts)
    s, FT = _sfx(kp.dtype)
    nf3 = list(nf) + [1] * (3 - dim)
    c = _cplx(c, kp.dtype)
    if fw is None:
        fw = np.zeros(nf3[::-1], c.dtype)
    xs = [np.ascontiguousarray(p, kp.dtype) for p in pts] + [None] * (3 - dim)
    fn = getattr(lib, "orc_spread" + s)
    fn.argtypes = [c_int, c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, FT, FT, c_int, c_void_p]
    fn(dim, xs[0].size, _p(xs[0]), _p(xs[1]), _p(xs[2]), _p(c), nf3[0], nf3[1], nf3[2], kp.ns, FT(kp.c), FT(kp.beta),
       kp.kerevalmeth, _p(fw))
    return fw


def interp(pts, fw, nf, kp):
    dim = len(pts)
    s, FT = _sfx(kp.dtype)
    nf3 = list(nf) + [1] * (3 - dim)
    fw = _cplx(fw, kp.dtype)
    xs = [np.ascontiguousarray(p, kp.dtype) for p in pts] + [None] * (3 - dim)
    c = np.zeros(xs[0].size, fw.dtype)
    fn = getattr(lib, "orc_interp" + s)
    fn.argtypes = [c_int, c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, FT, FT, c_int, c_void_p]
    fn(dim, xs[0].size, _p(xs[0]), _p(xs[1]), _p(xs[2]), _p(c), nf3[0], nf3[1], nf3[2], kp.ns, FT(kp.c), FT(kp.beta),
       kp.kerevalmeth, _p(fw))
    return c


def deconvolve(direction, fw, fk, modes, nf, kers, dtype):
    """direction 1: fk <- fw ; direction 2: fw <- fk (fw must be pre-zeroed)."""
    dim = len(modes)
    s, _ = _sfx(dtype)
    m = list(modes) + [1] * (3 - dim)
    n = list(nf) + [1] * (3 - dim)
    k = list(kers) + [None] * (3 - dim)
    fn = getattr(lib, "orc_deconvolve" + s)
    fn.argtypes = [c_int] * 8 + [c_void_p] * 5
    fn(direction, dim, m[0], m[1], m[2], n[0], n[1], n[2], _p(fw), _p(fk), _p(k[0]), _p(k[1]), _p(k[2]))


def plan_params(nufft_type, modes, tol, dtype, gpu_method=None, kerevalmeth=0, upsampfac=2.0):
    """Everything makeplan derives on the host: src/cufinufft.cu:78-273."""
    dim = len(modes)
    if gpu_method is None:
        gpu_method = 2 if nufft_type == 1 else 1        # src/cufinufft.cu:679-716
    kp = KernelParams(tol, dtype, upsampfac, kerevalmeth)
    bs, obs = default_binsize(dim, gpu_method)
    nf = [set_nf(modes[d], kp.ns, upsampfac, gpu_method, obs[d] if gpu_method == 4 else 1) for d in range(dim)]
    return kp, nf, bs[:dim], gpu_method


def nufft(nufft_type, modes, pts, data, tol, iflag=None, dtype=np.float64, kerevalmeth=0, upsampfac=2.0):
    """Full transform of ONE data vector through the oracle.
    modes = (ms[,mt[,mu]]) with x fastest; type 1: data=c[M] -> fk[mu][mt][ms];
    type 2: data=fk -> c[M].  Call stack: src/2d/cufinufft2d.cu:15-92 / :94-165."""
    dim = len(modes)
    if iflag is None:
        iflag = 1 if nufft_type == 1 else -1
    kp, nf, _, _ = plan_params(nufft_type, modes, tol, dtype, kerevalmeth=kerevalmeth, upsampfac=upsampfac)
    kers = [fwkerhalf(nf[d], kp) for d in range(dim)]
    cd = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
    fft = np.fft.ifftn if iflag >= 0 else np.fft.fftn      # cuFFT direction = iflag, unnormalised
    scale = float(np.prod(nf)) if iflag >= 0 else 1.0
    if nufft_type == 1:
        fw = spread(pts, data, nf, kp)
        fw = (fft(fw) * scale).astype(cd)
        fk = np.zeros(tuple(modes)[::-1], cd)
        deconvolve(1, fw, fk, modes, nf, kers, dtype)
        return fk
    fk = _cplx(data, dtype).reshape(tuple(modes)[::-1])
    fw = np.zeros(tuple(nf)[::-1], cd)
    deconvolve(2, fw, fk, modes, nf, kers, dtype)
    fw = (fft(fw) * scale).astype(cd)
    return interp(pts, fw, nf, kp)


def dirft1_sampled(pts, c, modes, iflag, modeidx):
    dim = len(pts)
    dtype = pts[0].dtype
    s, _ = _sfx(dtype)
    m = list(modes) + [1] * (3 - dim)
    xs = [np.ascontiguousarray(p) for p in pts] + [None] * (3 - dim)
    c = _cplx(c, dtype)
    idx = np.ascontiguousarray(modeidx, np.int64)
    out = np.zeros(idx.size, np.complex128)
    fn = getattr(lib, "orc_dirft1_sampled" + s)
    fn.argtypes = [c_int, c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    fn(dim, xs[0].size, _p(xs[0]), _p(xs[1]), _p(xs[2]), _p(c), iflag, m[0], m[1], m[2], idx.size, _p(idx), _p(out))
    return out


def dirft2_sampled(pts, fk, modes, iflag, ptidx):
    dim = len(pts)
    dtype = pts[0].dtype
    s, _ = _sfx(dtype)
    m = list(modes) + [1] * (3 - dim)
    xs = [np.ascontiguousarray(p) for p in pts] + [None] * (3 - dim)
    fk = _cplx(fk, dtype)
    idx = np.ascontiguousarray(ptidx, np.int64)
    out = np.zeros(idx.size, np.complex128)
    fn = getattr(lib, "orc_dirft2_sampled" + s)
    fn.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    fn(dim, _p(xs[0]), _p(xs[1]), _p(xs[2]), _p(fk), iflag, m[0], m[1], m[2], idx.size, _p(idx), _p(out))
    return out


def host_kernel(x, kp):
    """evaluate_kernel of contrib/spreadinterp.cpp:69-81 on an array."""
    s, FT = _sfx(kp.dtype)
    x = np.ascontiguousarray(x, kp.dtype)
    out = np.zeros(x.size, kp.dtype)
    fn = getattr(lib, "orc_host_kernel_vec" + s)
    fn.argtypes = [c_int, c_void_p, FT, FT, FT, c_void_p]
    fn(x.size, _p(x), FT(kp.beta), FT(kp.c), FT(kp.halfwidth), _p(out))
    return out


def horner_eval(w, x1, dtype):
    """ker[0..w-1] of eval_kernel_vec_Horner (src/cuspreadinterp.h:18-31) with OUR table."""
    s, FT = _sfx(dtype)
    out = np.zeros(16, dtype)
    fn = getattr(lib, "orc_horner_eval" + s)
    fn.argtypes = [c_int, FT, c_void_p]
    fn(int(w), FT(x1), _p(out))
    return out[:w]


def horner_table(w):
    lib.orc_horner_table_export.restype = ctypes.POINTER(c_double)
    lib.orc_horner_ncoef_export.restype = c_int
    nc = lib.orc_horner_ncoef_export(c_int(w))
    p = lib.orc_horner_table_export(c_int(w))
    flat = np.ctypeslib.as_array(p, shape=(18 * 16,)).copy()
    return flat.reshape(18, 16)[:nc, :w]
