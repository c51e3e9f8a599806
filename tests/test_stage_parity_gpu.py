"""Stage-level parity on the GPU (VERDICT r1 "missing 3", "weak 6"):
  * spread-only and interp-only grids against the REFERENCE's own CUSPREADnD / CUINTERPnD
    (src/cuspreadinterp.h:276-339; the reference's test/spread2d_test.cu:71-136 and
    interp2d_test.cu drive exactly these) on identical device inputs -- the kernel is
    un-normalised on both sides (SURVEY.md TL;DR 5), 1/2/3-D, both precisions, both evaluators,
    NUpts-driven and Subprob methods;
  * the same through the public switch gpu_spreadinterponly=1 (broken in v1.3, works here);
  * the host-pointer calls cufinufft[f]_setpts_host / _execute_host against the device-pointer path.
Tolerance: rel-l2 1e-5 (fp32) / 1e-12 (fp64), BASELINE.json north_star."""
import numpy as np
import pytest

import reflib
from helpers import cdtype, drop_exact_stencil_points, make_modes_data, make_points, make_strengths, rel_l2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float64: 1e-12}

STAGE_CASES = [
    # modes (x fastest), M, tol, dtype, dist, opts
    ((3000,), 100_000, 1e-5, np.float32, "uniform", dict(gpu_method=1)),
    ((3000,), 100_000, 1e-5, np.float32, "uniform", dict(gpu_method=2)),
    ((3000,), 100_000, 1e-11, np.float64, "uniform", dict(gpu_method=2, gpu_kerevalmeth=1)),
    ((256, 200), 400_000, 1e-3, np.float32, "uniform", dict(gpu_method=2)),
    ((256, 200), 400_000, 1e-5, np.float32, "cluster", dict(gpu_method=1)),
    ((256, 200), 400_000, 1e-5, np.float32, "cluster", dict(gpu_method=2, gpu_kerevalmeth=1)),
    ((256, 200), 200_000, 1e-9, np.float64, "uniform", dict(gpu_method=2)),
    ((256, 200), 200_000, 1e-9, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),
    ((128, 100), 60_000, 1e-12, np.float64, "wide", dict(gpu_method=2, gpu_kerevalmeth=1)),
    ((48, 40, 36), 300_000, 1e-5, np.float32, "cluster", dict(gpu_method=2)),
    ((48, 40, 36), 300_000, 1e-5, np.float32, "uniform", dict(gpu_method=1)),
    ((48, 40, 36), 300_000, 1e-4, np.float32, "uniform", dict(gpu_method=2, gpu_kerevalmeth=1)),
    ((32, 30, 28), 100_000, 1e-3, np.float64, "uniform", dict(gpu_method=2)),            # ns = 4: the reference's 16x16x2 tile fits its 48 KB
    ((32, 30, 28), 100_000, 1e-9, np.float64, "uniform", dict(gpu_method=2, gpu_binsizex=4, gpu_binsizey=4, gpu_binsizez=2)),   # ns = 10 needs small bins there
    ((32, 30, 28), 100_000, 1e-6, np.float64, "uniform", dict(gpu_method=1, gpu_kerevalmeth=1)),
    ((40, 30, 20), 100_000, 1e-3, np.float32, "onebin", dict(gpu_method=2)),     # reference spread3d_test's worst case
]


def _id(c):
    return "%s-M%d-%g-%s-%s-%s" % ("x".join(map(str, c[0])), c[1], c[2], np.dtype(c[3]).name, c[4],
                                   "_".join("%s%s" % kv for kv in c[5].items()))


def _setup(case, nufft_type):
    from cufinufft_b200 import gpuarray
    modes, M, tol, dtype, dist, opts = case
    dim = len(modes)
    pts = make_points(M, dim, dtype, seed=31, dist=dist)
    kp, nf, _, _ = orc.plan_params(nufft_type, modes, tol, dtype, gpu_method=opts.get("gpu_method"),
                                   kerevalmeth=opts.get("gpu_kerevalmeth", 0))
    pts = drop_exact_stencil_points(pts, nf, kp.ns)      # the reference reads an uninitialised weight there
    return pts, [gpuarray.to_gpu(p) for p in pts], nf


@pytest.mark.skipif(not reflib.available(), reason="reference library not built")
@pytest.mark.parametrize("case", STAGE_CASES, ids=_id)
def test_spread_only_vs_reference(case):
    from cufinufft_b200 import cufinufft, gpuarray
    modes, M, tol, dtype, dist, opts = case
    shape, cd = tuple(modes)[::-1], cdtype(dtype)
    pts, dev, nf = _setup(case, 1)
    M = pts[0].size
    gshape = tuple(nf)[::-1]
    c = gpuarray.to_gpu(make_strengths(M, dtype)[0])

    ref = reflib.RefPlan(1, modes, tol, dtype, **opts)
    ref.set_pts(dev)
    fw_ref = gpuarray.zeros(gshape, cd)
    ref.spread(c, fw_ref)
    want = fw_ref.get()
    ref.destroy()

    ours = cufinufft(1, shape, eps=tol, dtype=dtype, **opts)
    ours.set_pts(*dev[::-1])
    fw = gpuarray.empty(gshape, cd)
    ours.spread(c, fw)
    got = fw.get()
    ours.destroy()
    bound = TOL[dtype]
    if dist == "onebin":        # ~1e5 additions per cell in different orders on both sides: sqrt(M) eps noise floor
        bound = max(bound, 3 * np.sqrt(M) * np.finfo(dtype).eps)
    assert rel_l2(got, want) <= bound

    # the public switch: execute() of a gpu_spreadinterponly plan writes the fine grid into `fk`
    only = cufinufft(1, shape, eps=tol, dtype=dtype, gpu_spreadinterponly=1, **opts)
    only.set_pts(*dev[::-1])
    fw2 = gpuarray.empty(gshape, cd)
    only.execute(c, fw2)
    got2 = fw2.get()
    only.destroy()
    assert rel_l2(got2, want) <= bound


@pytest.mark.skipif(not reflib.available(), reason="reference library not built")
@pytest.mark.parametrize("case", STAGE_CASES, ids=_id)
def test_interp_only_vs_reference(case):
    from cufinufft_b200 import cufinufft, gpuarray
    modes, M, tol, dtype, dist, opts = case
    shape, cd = tuple(modes)[::-1], cdtype(dtype)
    pts, dev, nf = _setup(case, 2)
    M = pts[0].size
    fw = gpuarray.to_gpu(make_modes_data(nf, dtype, seed=9)[0])      # a random fine grid [nf3][nf2][nf1]

    # the reference's 1-D interpolation exists for method 1 only (src/1d/interp1d_wrapper.cu: "incorrect method,
    # should be 1"); the arithmetic is the same, so a 1-D method-2 request of ours is checked against its method 1
    ref_opts = dict(opts, gpu_method=1) if len(modes) == 1 else opts
    ref = reflib.RefPlan(2, modes, tol, dtype, **ref_opts)
    ref.set_pts(dev)
    c_ref = gpuarray.zeros((M,), cd)
    ref.interp(c_ref, fw)
    want = c_ref.get()
    ref.destroy()

    for engine in (0, 1, 2):                 # automatic, gather engine, tile engine
        ours = cufinufft(2, shape, eps=tol, dtype=dtype, **opts)
        ours.set_interp_engine(engine)
        ours.set_pts(*dev[::-1])
        c = gpuarray.zeros((M,), cd)
        ours.interp(c, fw)
        got = c.get()
        ours.destroy()
        assert rel_l2(got, want) <= TOL[dtype], engine

    only = cufinufft(2, shape, eps=tol, dtype=dtype, gpu_spreadinterponly=1, **opts)
    only.set_pts(*dev[::-1])
    c2 = gpuarray.zeros((M,), cd)
    only.execute(c2, fw)
    got2 = c2.get()
    only.destroy()
    assert rel_l2(got2, want) <= TOL[dtype]


HOST_CASES = [
    (1, (200, 150), 50_000, 1e-5, np.float32, 1, {}),
    (2, (200, 150), 50_000, 1e-5, np.float32, 3, {}),
    (1, (60, 50, 40), 80_000, 1e-9, np.float64, 2, {}),
    (2, (60, 50, 40), 80_000, 1e-9, np.float64, 1, {}),
    (2, (500,), 10_000, 1e-6, np.float32, 2, {}),
    (1, (64, 64), 0, 1e-5, np.float32, 1, {}),               # no points at all: zeros out, nothing read
    (2, (64, 64), 0, 1e-5, np.float32, 1, {}),
    (1, (64, 48), 30_000, 1e-5, np.float32, 2, dict(gpu_spreadinterponly=1)),
    # single transforms of >= 2e6 points take the chunked copy/compute pipeline of execute_host (csrc/plan.cu:
    # PIPE_MIN_POINTS): 4 chunks of the caller's index range for type 1, 8 for type 2
    (1, (256, 200), 2_500_000, 1e-5, np.float32, 1, {}),
    (2, (256, 200), 2_500_000, 1e-9, np.float64, 1, {}),
    (1, (48, 40, 36), 2_100_000, 1e-6, np.float64, 1, {}),
    (2, (48, 40, 36), 2_100_003, 1e-5, np.float32, 1, {}),
]


@pytest.mark.parametrize("case", HOST_CASES, ids=lambda c: "t%d-%s-M%d-%s-n%d%s" % (
    c[0], "x".join(map(str, c[1])), c[2], np.dtype(c[4]).name, c[5], "-only" if c[6] else ""))
def test_host_pointer_calls_match_device_path(case):
    """setpts_host + execute_host (what bench.py's e2e leg times) == set_pts + execute on device arrays, bit for bit
    for type 2 (no atomics) and to accumulation-order noise for type 1."""
    from cufinufft_b200 import cufinufft, gpuarray
    nufft_type, modes, M, tol, dtype, ntransf, opts = case
    dim, shape, cd = len(modes), tuple(modes)[::-1], cdtype(dtype)
    pts = make_points(M, dim, dtype, seed=5)
    only = bool(opts.get("gpu_spreadinterponly"))

    dplan = cufinufft(nufft_type, shape, n_trans=ntransf, eps=tol, dtype=dtype, **opts)
    oshape = tuple(dplan.geometry()["nf%d" % (d + 1)] for d in range(dim))[::-1] if only else shape
    c_in = make_strengths(max(M, 1), dtype, ntransf=ntransf)[:, :M]
    fk_in = make_modes_data(oshape[::-1], dtype, ntransf=ntransf)
    dev = [gpuarray.to_gpu(p) for p in pts]
    dplan.set_pts(*dev[::-1])
    if nufft_type == 1:
        cg, fg = gpuarray.to_gpu(np.ascontiguousarray(c_in)), gpuarray.zeros((ntransf,) + oshape, cd)
        dplan.execute(cg, fg)
        want = fg.get()
    else:
        cg, fg = gpuarray.zeros((ntransf, max(M, 1)), cd), gpuarray.to_gpu(fk_in)
        dplan.execute(cg, fg)
        want = cg.get()[:, :M]
    dplan.destroy()

    hplan = cufinufft(nufft_type, shape, n_trans=ntransf, eps=tol, dtype=dtype, **opts)
    hplan.set_pts_host(*pts[::-1])
    if nufft_type == 1:
        c_h = np.ascontiguousarray(c_in)
        fk_h = np.full((ntransf,) + oshape, np.nan + 0j, cd)
        hplan.execute_host(c_h, fk_h)
        got = fk_h
    else:
        c_h = np.full((ntransf, M), np.nan + 0j, cd)
        fk_h = np.ascontiguousarray(fk_in)
        hplan.execute_host(c_h, fk_h)
        got = c_h
    hplan.destroy()
    assert np.all(np.isfinite(got.view(dtype)))
    if M == 0:
        assert not got.any()
    elif nufft_type == 2:
        assert np.array_equal(got, want)
    else:
        assert rel_l2(got, want) <= 10 * np.finfo(dtype).eps
