"""GPU parity against the REFERENCE ITSELF: cuFINUFFT v1.3 built for sm_100 from
/root/reference (oracle/_ref/libcufinufft_ref.so) and ours run in the same process on
identical device buffers.  Gates (BASELINE.json north_star): bin counts / offsets / subproblem
map bit-exact; outputs rel-l2 <= 1e-5 (fp32) / 1e-12 (fp64)."""
import numpy as np
import pytest

import reflib
from helpers import cdtype, drop_exact_stencil_points, make_modes_data, make_points, make_strengths, rel_l2
from oracle import oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reflib.available(), reason="reference library not built")]

TOL = {np.float32: 1e-5, np.float64: 1e-12}

CASES = [
    # type, modes (x fastest), M, tol, dtype, dist, opts
    (1, (1000, 1000), 1_000_000, 1e-3, np.float32, "uniform", dict(gpu_method=2)),      # config 1 at M/10
    (2, (512, 512), 1_000_000, 1e-9, np.float64, "uniform", dict(gpu_method=1, gpu_sort=1)),   # config 2 scaled
    (1, (64, 64, 64), 1_000_000, 1e-5, np.float32, "cluster", dict(gpu_method=2)),      # config 3 scaled
    (2, (64, 64, 64), 300_000, 1e-9, np.float64, "uniform", dict(gpu_method=1, gpu_sort=1)),   # config 5 scaled
    (1, (512, 512), 262_144, 1e-4, np.float32, "uniform", dict(gpu_method=2)),          # config 4 shape
    (2, (512, 512), 262_144, 1e-4, np.float32, "uniform", dict(gpu_method=1)),
    (1, (200, 150), 50_000, 1e-6, np.float32, "wide", dict(gpu_method=1)),
    (1, (200, 150), 50_000, 1e-12, np.float64, "wide", dict(gpu_method=2)),
    (2, (200, 150), 50_000, 1e-6, np.float32, "wide", dict(gpu_method=2)),
    (1, (3000,), 200_000, 1e-6, np.float32, "uniform", dict(gpu_method=2)),
    (2, (3000,), 200_000, 1e-12, np.float64, "uniform", dict(gpu_method=1)),
    (1, (40, 30, 20), 100_000, 1e-3, np.float32, "onebin", dict(gpu_method=2)),
    (1, (8, 8), 1000, 1e-3, np.float32, "uniform", dict(gpu_method=2)),
    (1, (64, 48), 20000, 1e-4, np.float32, "uniform", dict(gpu_method=2, gpu_kerevalmeth=1)),
    (2, (64, 48), 20000, 1e-4, np.float32, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),
    (1, (24, 20, 16), 20000, 1e-5, np.float32, "uniform", dict(gpu_method=2, gpu_kerevalmeth=1)),
    # fp64 Horner (gpu_kerevalmeth=1) with the reference's own coefficient table: w = 5, 7, 10, 13 (VERDICT r1 "weak 2")
    (1, (64, 48), 20000, 1e-4, np.float64, "uniform", dict(gpu_method=2, gpu_kerevalmeth=1)),
    (2, (64, 48), 20000, 1e-6, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),
    (1, (64, 48), 20000, 1e-9, np.float64, "cluster", dict(gpu_method=2, gpu_kerevalmeth=1)),
    (2, (64, 48), 20000, 1e-12, np.float64, "uniform", dict(gpu_method=2, gpu_kerevalmeth=1)),
    (1, (24, 20, 16), 20000, 1e-6, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),   # (method 2: the reference's tile exceeds its 48 KB at w = 7)
    (2, (24, 20, 16), 20000, 1e-9, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),
    (1, (24, 20, 16), 20000, 1e-2, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),     # w = 3
    (2, (500,), 20000, 1e-1, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),           # w = 2
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "t%d-%s-M%d-%g-%s-%s-%s" % (
    c[0], "x".join(map(str, c[1])), c[2], c[3], np.dtype(c[4]).name, c[5], "_".join("%s%s" % kv for kv in c[6].items())))
def test_against_reference_library(case):
    from cufinufft_b200 import cufinufft, gpuarray
    nufft_type, modes, M, tol, dtype, dist, opts = case
    dim = len(modes)
    shape = tuple(modes)[::-1]
    cd = cdtype(dtype)
    pts = make_points(M, dim, dtype, seed=21, dist=dist)
    kp, nf, _, _ = orc.plan_params(nufft_type, modes, tol, dtype, gpu_method=opts.get("gpu_method"),
                                   kerevalmeth=opts.get("gpu_kerevalmeth", 0))
    pts = drop_exact_stencil_points(pts, nf, kp.ns)      # the reference reads an uninitialised weight there
    M = pts[0].size
    dev = [gpuarray.to_gpu(p) for p in pts]

    ours = cufinufft(nufft_type, shape, eps=tol, dtype=dtype, **opts)
    ours.set_pts(*dev[::-1])
    ref = reflib.RefPlan(nufft_type, modes, tol, dtype, **opts)
    ref.set_pts(dev)

    # plan-time numbers
    go, gr = ours.geometry(), ref.geometry()
    for k in ("nf1", "nf2", "nf3", "ns", "nbins1", "nbins2", "nbins3", "binsx", "binsy", "binsz"):
        assert go[k] == gr[k], k
    for d in range(dim):
        po, pr = ours.phihat(d), ref.phihat(d)
        used = modes[d] // 2 + 1
        assert np.max(np.abs(po[:used] - pr[:used]) / np.abs(pr[:used])) <= (2e-7 if dtype == np.float32 else 1e-13)

    # bin sort: bit-exact counts / offsets / subproblem map; idxnupts equal as per-bin sets
    if opts.get("gpu_method") == 2 or opts.get("gpu_sort", 1):
        lo, lr = ours.bin_layout(), ref.bin_layout()
        assert np.array_equal(lo["binsize"], lr["binsize"])
        assert np.array_equal(lo["binstartpts"], lr["binstartpts"])
        if opts.get("gpu_method") == 2:
            for k in ("numsubprob", "subprobstartpts", "subprob_to_bin"):
                assert np.array_equal(lo[k], lr[k]), k
            assert lo["totalnumsubprob"] == lr["totalnumsubprob"]
        starts, sizes = lr["binstartpts"], lr["binsize"]
        assert np.array_equal(np.sort(lo["idxnupts"]), np.arange(M))
        for b in np.flatnonzero(sizes)[:: max(1, len(sizes) // 300)]:
            s, n = starts[b], sizes[b]
            assert np.array_equal(np.sort(lo["idxnupts"][s:s + n]), np.sort(lr["idxnupts"][s:s + n]))

    # full transform on identical device inputs
    if nufft_type == 1:
        c = gpuarray.to_gpu(make_strengths(M, dtype)[0])
        fo, fr = gpuarray.zeros(shape, cd), gpuarray.zeros(shape, cd)
        ours.execute(c, fo)
        ref.execute(c, fr)
        a, b = fo.get(), fr.get()
    else:
        fk = gpuarray.to_gpu(make_modes_data(modes, dtype)[0])
        co, cr = gpuarray.zeros((M,), cd), gpuarray.zeros((M,), cd)
        ours.execute(co, fk)
        ref.execute(cr, fk)
        a, b = co.get(), cr.get()
    err = rel_l2(a, b)
    ref.destroy()
    ours.destroy()
    assert err <= TOL[dtype], err
