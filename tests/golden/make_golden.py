#!/usr/bin/env python
"""tests/golden/make_golden.py -- regenerate tests/golden/host_math.npz from the REFERENCE's
own host code (oracle/_ref/libref_host.so = /root/reference/contrib/*.cpp compiled by
oracle/Makefile target ref_host; only possible where /root/reference exists).

The reference stores no golden vectors for this path, so these outputs of the
reference itself are what pins the oracle: kernel parameters, fine-grid sizes, Gauss-Legendre
nodes, the phihat quadrature precomputation and CPU phihat, host kernel values, the generated
Horner table evaluated at sample offsets, and dirft2d direct sums.
    make -C oracle ref_host && python tests/golden/make_golden.py
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_host.so"))

TOLS = [0.5, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10, 1e-12, 1e-14, 1e-17]
NF_NS = [(2000, 4), (4096, 10), (512, 6), (1024, 5), (1024, 10), (16, 4), (96, 7), (400, 2), (6000, 7), (90, 16), (50, 3)]


def p(a):
    return a.ctypes.data_as(c_void_p)


def main():
    out = {}
    rows = []
    for sfx, FT, dt in (("", c_double, np.float64), ("f", c_float, np.float32)):
        fn = getattr(L, "refh_setup_spreader" + sfx)
        fn.argtypes = [FT, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        fn.restype = c_int
        for tol in TOLS:
            ns = c_int()
            b, h, c = FT(), FT(), FT()
            ier = fn(FT(tol), 2.0, 0, ctypes.byref(ns), ctypes.byref(b), ctypes.byref(h), ctypes.byref(c))
            rows.append((0 if sfx == "" else 1, tol, ier, ns.value, b.value, h.value, c.value))
    out["setup_spreader"] = np.array(rows, np.float64)       # float values are exactly representable in double

    L.refh_next235beven.restype = c_int
    ns_in = [1, 2, 3, 7, 16, 31, 100, 121, 1000, 1001, 2000, 2047, 4096, 4097, 12345, 99991, 1000003]
    out["next235_in"] = np.array([(n, b) for n in ns_in for b in (1, 4, 8)], np.int64)
    out["next235_out"] = np.array([L.refh_next235beven(int(n), int(b)) for n, b in out["next235_in"]], np.int64)
    L.refh_set_nf.restype = c_int
    L.refh_set_nf.argtypes = [c_int, c_double, c_int, c_int, c_int]
    snf = [(ms, ns, meth, 8 if meth == 4 else 1) for ms in (1, 5, 8, 100, 256, 512, 1000, 2048) for ns in (2, 4, 6, 10, 16)
           for meth in (1, 2, 4)]
    out["set_nf_in"] = np.array(snf, np.int64)
    out["set_nf_out"] = np.array([L.refh_set_nf(ms, 2.0, ns, meth, ob) for ms, ns, meth, ob in snf], np.int64)

    for n in (8, 16, 22, 34, 52):
        x, w = np.zeros(n), np.zeros(n)
        L.refh_legendre(c_int(n), p(x), p(w))
        out["legendre_x_%d" % n], out["legendre_w_%d" % n] = x, w

    for sfx, FT, dt in (("", c_double, np.float64), ("f", c_float, np.float32)):
        ss = getattr(L, "refh_setup_spreader" + sfx)
        pre = getattr(L, "refh_fseries_precomp" + sfx)
        pre.argtypes = [c_int, c_int, FT, FT, FT, c_void_p, c_void_p]
        cpu = getattr(L, "refh_fseries_cpu" + sfx)
        cpu.argtypes = [c_int, c_int, FT, FT, FT, c_void_p]
        ek = getattr(L, "refh_evaluate_kernel" + sfx)
        ek.argtypes = [FT, c_int, FT, FT, FT]
        ek.restype = FT
        hz = getattr(L, "refh_horner" + sfx)
        hz.argtypes = [c_int, FT, c_void_p]
        tol_of_ns = {2: 1e-1, 3: 1e-2, 4: 1e-3, 5: 1e-4, 6: 1e-5, 7: 1e-6, 8: 1e-7, 9: 1e-8, 10: 1e-9, 16: 1e-15}
        for nf, ns in NF_NS:
            if dt == np.float32 and ns > 8:
                continue
            nsv = c_int()
            b, h, c = FT(), FT(), FT()
            ss(FT(tol_of_ns[ns]), 2.0, 0, ctypes.byref(nsv), ctypes.byref(b), ctypes.byref(h), ctypes.byref(c))
            assert nsv.value == ns, (ns, nsv.value)
            q = int(2 + 3.0 * (ns / 2.0))
            f, a = np.zeros(q, dt), np.zeros(2 * q, np.float64)
            pre(nf, ns, b, h, c, p(f), p(a))
            ker = np.zeros(nf // 2 + 1, dt)
            cpu(nf, ns, b, h, c, p(ker))
            key = "%s_nf%d_ns%d" % (np.dtype(dt).name, nf, ns)
            out["fser_f_" + key], out["fser_a_" + key], out["fwkerhalf_cpu_" + key] = f, a, ker
        xs = np.linspace(-8.5, 8.5, 137).astype(dt)
        for ns in (2, 4, 5, 6, 7, 10, 16):
            nsv = c_int()
            b, h, c = FT(), FT(), FT()
            ss(FT(tol_of_ns[ns] if not (dt == np.float32 and ns > 8) else 1e-3), 2.0, 0, ctypes.byref(nsv), ctypes.byref(b),
               ctypes.byref(h), ctypes.byref(c))
            if nsv.value != ns:
                continue
            out["evalker_%s_ns%d" % (np.dtype(dt).name, ns)] = np.array([ek(FT(x), ns, b, h, c) for x in xs], dt)
        out["evalker_x_" + np.dtype(dt).name] = xs
        # Horner table: ker[0..w-1] at 9 offsets x1 in [-w/2, -w/2+1]
        for w in range(2, 17):
            offs = (-w / 2.0 + np.linspace(0.0, 1.0, 9)).astype(dt)
            vals = np.zeros((9, w), dt)
            for i, x1 in enumerate(offs):
                row = np.zeros(16, dt)
                hz(w, FT(x1), p(row))
                vals[i] = row[:w]
            out["horner_%s_w%d" % (np.dtype(dt).name, w)] = vals

        # direct sums (contrib/dirft2d.cpp)
        rng = np.random.default_rng(3)
        nj, ms, mt = 60, 12, 9
        x = rng.uniform(-np.pi, np.pi, nj).astype(dt)
        y = rng.uniform(-np.pi, np.pi, nj).astype(dt)
        cd = np.complex128 if dt == np.float64 else np.complex64
        c = (rng.uniform(-1, 1, nj) + 1j * rng.uniform(-1, 1, nj)).astype(cd)
        f1 = np.zeros(ms * mt, cd)
        d1 = getattr(L, "refh_dirft2d1" + sfx)
        d1.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        d1(nj, p(x), p(y), p(c), 1, ms, mt, p(f1))
        fk = (rng.uniform(-1, 1, ms * mt) + 1j * rng.uniform(-1, 1, ms * mt)).astype(cd)
        c2 = np.zeros(nj, cd)
        d2 = getattr(L, "refh_dirft2d2" + sfx)
        d2.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
        d2(nj, p(x), p(y), p(c2), -1, ms, mt, p(fk))
        n = np.dtype(dt).name
        out["dirft_x_" + n], out["dirft_y_" + n], out["dirft_c_" + n] = x, y, c
        out["dirft_f1_" + n], out["dirft_fk_" + n], out["dirft_c2_" + n] = f1, fk, c2
    out["dirft_shape"] = np.array([60, 12, 9])
    np.savez_compressed(os.path.join(HERE, "host_math.npz"), **out)
    print("wrote host_math.npz with %d arrays" % len(out))


if __name__ == "__main__":
    main()
