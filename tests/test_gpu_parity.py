"""GPU parity tests: our CUDA path (through the C ABI / Python class) against the CPU
oracle on identical seeded inputs.  Tolerances: BASELINE.json north_star -- bin counts and
offsets bit-exact; transforms rel-l2 <= 1e-5 (fp32) / 1e-12 (fp64) against the reference
arithmetic; outputs within the requested tol of direct sums."""
import numpy as np
import pytest

from helpers import cdtype, gpu_nufft, make_modes_data, make_points, make_strengths, rel_l2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_PARITY = {np.float32: 1e-5, np.float64: 1e-12}


def _check_bins(plan, pts, dtype, maxsub=1024):
    lay = plan.bin_layout()
    dim = len(pts)
    nf = [lay["nf1"], lay["nf2"], lay["nf3"]][:dim]
    bs = [lay["binsx"], lay["binsy"], lay["binsz"]][:dim]
    ref = orc.binsort(pts, nf, bs, maxsub)
    for key in ("binsize", "binstartpts", "numsubprob", "subprobstartpts", "subprob_to_bin"):
        assert np.array_equal(lay[key], ref[key]), key
    assert lay["totalnumsubprob"] == ref["totalnumsubprob"]
    # idxnupts: a permutation, bin-major, equal to the oracle's as per-bin sets
    idx, M = lay["idxnupts"], pts[0].size
    assert np.array_equal(np.sort(idx), np.arange(M))
    starts = ref["binstartpts"]
    ours = np.sort(idx[: M]) if M == 0 else None
    for b in np.flatnonzero(ref["binsize"])[:: max(1, len(starts) // 200)]:
        s, n = starts[b], ref["binsize"][b]
        assert np.array_equal(np.sort(idx[s:s + n]), np.sort(ref["idxnupts"][s:s + n]))


CASES = [
    # (type, modes, M, tol, dtype, opts)
    (1, (64, 48), 20000, 1e-4, np.float32, {}),
    (2, (64, 48), 20000, 1e-4, np.float32, {}),
    (1, (100, 80), 30000, 1e-3, np.float32, {}),                       # config-1 shape (ns=4), nf not a bin multiple
    (1, (40, 36), 5000, 1e-6, np.float32, {}),
    (1, (64, 48), 20000, 1e-9, np.float64, {}),
    (2, (64, 48), 20000, 1e-9, np.float64, {}),                        # config-2 shape (ns=10, GM-sort interp)
    (1, (64, 48), 20000, 1e-4, np.float32, dict(gpu_method=1)),
    (1, (64, 48), 20000, 1e-4, np.float32, dict(gpu_method=1, gpu_sort=0)),
    (2, (64, 48), 20000, 1e-4, np.float32, dict(gpu_method=2)),
    (1, (24, 20, 16), 20000, 1e-5, np.float32, {}),                    # config-3 shape (ns=6, SM 3-D)
    (2, (24, 20, 16), 20000, 1e-5, np.float32, {}),
    (1, (24, 20, 16), 20000, 1e-5, np.float32, dict(gpu_method=4)),
    (1, (20, 18, 16), 8000, 1e-9, np.float64, {}),
    (2, (20, 18, 16), 8000, 1e-9, np.float64, {}),                     # config-5 shape (ns=10, 3-D fp64)
    (1, (200,), 5000, 1e-5, np.float32, {}),
    (2, (200,), 5000, 1e-5, np.float32, {}),
    (1, (200,), 5000, 1e-10, np.float64, {}),
    (2, (200,), 5000, 1e-10, np.float64, {}),
    (1, (8, 8), 100, 1e-3, np.float32, {}),                            # nf=16 < bin 32 ("make check" 8x8 cases)
    (1, (8, 8, 8), 32, 1e-3, np.float32, dict(gpu_sort=0, gpu_maxsubprobsize=10)),   # reference test_opts
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "t%d-%s-M%d-%g-%s-%s" % (
    c[0], "x".join(map(str, c[1])), c[2], c[3], np.dtype(c[4]).name, "_".join("%s%s" % kv for kv in c[5].items())))
def test_transform_vs_oracle(case):
    nufft_type, modes, M, tol, dtype, opts = case
    dim = len(modes)
    pts = make_points(M, dim, dtype, seed=10 + dim)
    if nufft_type == 1:
        data = make_strengths(M, dtype)
    else:
        data = make_modes_data(modes, dtype)
    out, plan = gpu_nufft(nufft_type, modes, pts, data, tol, dtype, return_plan=True, **opts)
    ref = orc.nufft(nufft_type, modes, pts, data[0], tol, dtype=dtype)
    err = rel_l2(out[0], ref)
    assert err <= TOL_PARITY[dtype], err
    if plan.geometry()["method"] == 2 or opts.get("gpu_sort", 1):
        if opts.get("gpu_method", 0) != 4:
            _check_bins(plan, pts, dtype, opts.get("gpu_maxsubprobsize", 1024))
    # accuracy against the direct sum at sampled outputs (<= requested tol, with the usual
    # factor for the l2 -> sampled max conversion)
    rng = np.random.default_rng(5)
    if nufft_type == 1:
        idx = rng.integers(0, int(np.prod(modes)), 50)
        exact = orc.dirft1_sampled(pts, data[0], modes, 1, idx)
        got = out[0].ravel()[idx]
        scale = np.abs(exact).max()
    else:
        idx = rng.integers(0, M, 50)
        exact = orc.dirft2_sampled(pts, data[0], modes, -1, idx)
        got = out[0][idx]
        scale = np.abs(exact).max()
    floor = 3e-6 if dtype == np.float32 else 1e-13
    assert np.abs(got - exact).max() / scale <= max(10 * tol, floor)


# Inputs dense enough for setpts to split the reference's bins into internal sub-bins (one or more
# halvings, csrc/spread.cu: choose_internal_bins) and distributions that exercise both the
# run-merged and the point-by-point paths of the SM spread kernel.  The reference-facing bin arrays
# must stay bit-exact, the transform within the parity tolerance.
SUBBIN_CASES = [
    # (modes, M, tol, dtype, dist, opts)
    ((128, 96), 400000, 1e-3, np.float32, "uniform", {}),
    ((100, 80), 300000, 1e-4, np.float32, "cluster", {}),                     # nf not a multiple of the bins
    ((64, 64), 60000, 1e-4, np.float32, "onebin", {}),                        # every point in one bin: long runs
    ((32, 32, 32), 300000, 1e-5, np.float32, "uniform", {}),
    ((24, 20, 16), 200000, 1e-5, np.float32, "cluster", dict(gpu_maxsubprobsize=300)),
    ((20, 18, 16), 150000, 1e-9, np.float64, "uniform", {}),                  # wide fp64 stencil: MERGE = false path
    ((48, 40), 200000, 1e-9, np.float64, "cluster", {}),
]


@pytest.mark.parametrize("case", SUBBIN_CASES, ids=lambda c: "%s-M%d-%s-%s" % ("x".join(map(str, c[0])), c[1], np.dtype(c[3]).name, c[4]))
def test_type1_with_internal_subbins(case):
    modes, M, tol, dtype, dist, opts = case
    dim = len(modes)
    pts = make_points(M, dim, dtype, seed=77, dist=dist)
    data = make_strengths(M, dtype)
    out, plan = gpu_nufft(1, modes, pts, data, tol, dtype, return_plan=True, **opts)
    ref = orc.nufft(1, modes, pts, data[0], tol, dtype=dtype)
    err = rel_l2(out[0], ref)
    # every point in the same few cells: 6e4 fp32 additions per cell in an order that differs between
    # the two implementations (and between runs of either): sqrt(N) * eps ~ 1.5e-5 is the noise floor
    assert err <= TOL_PARITY[dtype] * (4 if dist == "onebin" else 1), err
    _check_bins(plan, pts, dtype, opts.get("gpu_maxsubprobsize", 1024))
    # a second setpts with few points on the same plan falls back to the reference's bins
    few = [p[:500].copy() for p in pts]
    from cufinufft_b200 import gpuarray
    dev = [gpuarray.to_gpu(p) for p in few]
    plan.set_pts(*dev[::-1])
    cg = gpuarray.to_gpu(np.ascontiguousarray(data[:, :500]))
    fkg = gpuarray.zeros((1,) + tuple(modes)[::-1], cdtype(dtype))
    plan.execute(cg, fkg)
    ref2 = orc.nufft(1, modes, few, data[0, :500], tol, dtype=dtype)
    assert rel_l2(fkg.get()[0], ref2) <= TOL_PARITY[dtype]
    _check_bins(plan, few, dtype, opts.get("gpu_maxsubprobsize", 1024))


# The two-level point order (bins globally, stencil cells per work item: csrc/setpts.cu
# local_sort_kernel) is chosen automatically only for histograms beyond 2 GB; forced here on
# small inputs of every kind.  Same results, same reference-facing bin arrays.
TWO_LEVEL_CASES = [
    (1, (128, 96), 400000, 1e-3, np.float32, "uniform", {}),
    (1, (64, 64), 60000, 1e-4, np.float32, "onebin", {}),
    (2, (64, 48), 50000, 1e-9, np.float64, "uniform", {}),
    (1, (32, 32, 32), 300000, 1e-5, np.float32, "cluster", {}),
    (2, (20, 18, 16), 100000, 1e-9, np.float64, "wide", {}),
    (1, (24, 20, 16), 150000, 1e-6, np.float32, "cluster", dict(gpu_maxsubprobsize=3000)),
    (1, (24, 20, 16), 50000, 1e-5, np.float32, "uniform", dict(gpu_method=1)),
    (1, (300,), 100000, 1e-5, np.float32, "uniform", {}),
]


@pytest.mark.parametrize("case", TWO_LEVEL_CASES, ids=lambda c: "t%d-%s-M%d-%s-%s" % (c[0], "x".join(map(str, c[1])), c[2], np.dtype(c[4]).name, c[5]))
def test_two_level_point_order(case):
    nufft_type, modes, M, tol, dtype, dist, opts = case
    dim = len(modes)
    pts = make_points(M, dim, dtype, seed=91, dist=dist)
    data = make_strengths(M, dtype) if nufft_type == 1 else make_modes_data(modes, dtype)
    out, plan = gpu_nufft(nufft_type, modes, pts, data, tol, dtype, return_plan=True, sort_levels=2, **opts)
    ref = orc.nufft(nufft_type, modes, pts, data[0], tol, dtype=dtype)
    slack = 4 if dist == "onebin" else 1                 # fp32 accumulation-order noise, see above
    assert rel_l2(out[0], ref) <= TOL_PARITY[dtype] * slack
    one = gpu_nufft(nufft_type, modes, pts, data, tol, dtype, sort_levels=1, **opts)
    assert rel_l2(out[0], one[0]) <= TOL_PARITY[dtype] * slack
    _check_bins(plan, pts, dtype, opts.get("gpu_maxsubprobsize", 1024))
    assert plan.launch_counts()["setpts"] >= 9          # the local sort ran


# The coarse partition in front of the counting sort (csrc/setpts.cu: coarse_count / coarse_scatter) switches itself
# on only for point sets beyond a quarter of L2; forced here (sort_levels + 8) on small inputs of every kind, with
# the one-level and the two-level order behind it, against the direct scatter (+ 4).  Same results (the order of
# the points inside a key is free), same reference-facing bin arrays.
@pytest.mark.parametrize("levels", [1, 2])
@pytest.mark.parametrize("case", TWO_LEVEL_CASES, ids=lambda c: "t%d-%s-M%d-%s-%s" % (c[0], "x".join(map(str, c[1])), c[2], np.dtype(c[4]).name, c[5]))
def test_coarse_partitioned_sort(case, levels):
    nufft_type, modes, M, tol, dtype, dist, opts = case
    dim = len(modes)
    pts = make_points(M, dim, dtype, seed=92, dist=dist)
    data = make_strengths(M, dtype) if nufft_type == 1 else make_modes_data(modes, dtype)
    out, plan = gpu_nufft(nufft_type, modes, pts, data, tol, dtype, return_plan=True, sort_levels=8 + levels, **opts)
    direct, plan0 = gpu_nufft(nufft_type, modes, pts, data, tol, dtype, return_plan=True, sort_levels=4 + levels, **opts)
    slack = 4 if dist == "onebin" else 1                 # fp32 accumulation-order noise, see above
    assert rel_l2(out[0], direct[0]) <= TOL_PARITY[dtype] * slack
    ref = orc.nufft(nufft_type, modes, pts, data[0], tol, dtype=dtype)
    assert rel_l2(out[0], ref) <= TOL_PARITY[dtype] * slack
    _check_bins(plan, pts, dtype, opts.get("gpu_maxsubprobsize", 1024))
    assert plan.launch_counts()["setpts"] == plan0.launch_counts()["setpts"] + 5      # count, 3-phase scan, scatter ran


# Double precision, 3-D, stencils of nine points and more with gpu_method 2: the plane-owner spreading engine
# (csrc/spread_plane.cuh: one block-shared tile per reference bin, a warp per z plane).  The reference refuses these
# plans (its tile exceeds 48 KB of shared memory), so the truth is the oracle and our own GM engine.
PLANE_CASES = [
    ((20, 18, 16), 30_000, 1e-9, "uniform", {}),                                   # ns = 10
    ((24, 20, 16), 200_000, 1e-9, "cluster", {}),                                  # several batches and work items per bin
    ((16, 16, 16), 60_000, 1e-8, "onebin", {}),                                    # ns = 9 (the conflict-free stride does not fit: natural stride)
    ((16, 14, 12), 20_000, 1e-10, "wide", {}),                                     # ns = 11
    ((20, 18, 16), 30_000, 1e-9, "uniform", dict(gpu_kerevalmeth=1)),              # Horner
    ((20, 18, 16), 30_000, 1e-9, "uniform", dict(gpu_binsizex=8, gpu_binsizey=12, gpu_binsizez=3)),   # other bins: 13 tile planes
    ((20, 18, 16), 50_000, 1e-9, "cluster", dict(gpu_maxsubprobsize=300)),
]


@pytest.mark.parametrize("case", PLANE_CASES, ids=lambda c: "%s-M%d-%g-%s%s" % ("x".join(map(str, c[0])), c[1], c[2], c[3], "".join("-%s%s" % kv for kv in c[4].items())))
def test_plane_owner_spread_engine(case):
    modes, M, tol, dist, opts = case
    dtype = np.float64
    pts = make_points(M, 3, dtype, seed=17, dist=dist)
    data = make_strengths(M, dtype, ntransf=2)
    out, plan = gpu_nufft(1, modes, pts, data, tol, dtype, ntransf=2, maxbatch=2, return_plan=True, gpu_method=2, **opts)
    slack = 40 if dist == "onebin" else 1                   # sqrt(M) * eps accumulation-order noise when every point shares a few cells
    for t in range(2):
        ref = orc.nufft(1, modes, pts, data[t], tol, dtype=dtype, kerevalmeth=opts.get("gpu_kerevalmeth", 0))
        assert rel_l2(out[t], ref) <= TOL_PARITY[dtype] * slack
    gm = gpu_nufft(1, modes, pts, data, tol, dtype, ntransf=2, maxbatch=2, **dict(opts, gpu_method=1))
    assert rel_l2(out, gm) <= TOL_PARITY[dtype] * slack
    _check_bins(plan, pts, dtype, opts.get("gpu_maxsubprobsize", 1024))


# Nonstandard upsampling factors (opts.upsampfac != 2: kernel width and beta from the cutoff formulas of
# contrib/spreadinterp.cpp:43-62, fine grid sigma * modes; direct kernel evaluation only).  The reference
# accepts them through the same opts field; sigma = 1.25 shrinks the fine grid of a 3-D transform 4x.
@pytest.mark.parametrize("case", [
    (1, (60, 50), 30000, 1e-4, np.float32, 1.25), (2, (60, 50), 30000, 1e-4, np.float32, 1.25),
    (1, (40, 36), 20000, 1e-8, np.float64, 1.5), (2, (20, 18, 16), 20000, 1e-6, np.float64, 1.25),
    (1, (24, 20, 16), 20000, 1e-3, np.float32, 3.0),
    # single precision with kernels wider than upsampfac 2 ever needs (ns = 10, 13, 12):
    (1, (64, 48), 20000, 1e-6, np.float32, 1.25), (2, (64, 48), 20000, 3e-7, np.float32, 1.25),
    (1, (16, 14, 12), 5000, 1e-6, np.float32, 1.3), (2, (300,), 5000, 1e-6, np.float32, 1.25),
], ids=lambda c: "t%d-%s-%g-%s-sigma%g" % (c[0], "x".join(map(str, c[1])), c[3], np.dtype(c[4]).name, c[5]))
def test_nonstandard_upsampfac(case):
    nufft_type, modes, M, tol, dtype, sigma = case
    dim = len(modes)
    pts = make_points(M, dim, dtype, seed=12)
    data = make_strengths(M, dtype) if nufft_type == 1 else make_modes_data(modes, dtype)
    out, plan = gpu_nufft(nufft_type, modes, pts, data, tol, dtype, return_plan=True, upsampfac=sigma)
    g = plan.geometry()
    assert sigma * modes[0] - 1 <= g["nf1"] <= 1.2 * max(sigma * modes[0], 2 * g["ns"]) + 8      # next 2^a 3^b 5^c even above both
    ref = orc.nufft(nufft_type, modes, pts, data[0], tol, dtype=dtype, upsampfac=sigma)
    rng = np.random.default_rng(5)
    if nufft_type == 1:
        idx = rng.integers(0, int(np.prod(modes)), 40)
        exact, got, want = orc.dirft1_sampled(pts, data[0], modes, 1, idx), out[0].ravel()[idx], ref.ravel()[idx]
    else:
        idx = rng.integers(0, M, 40)
        exact, got, want = orc.dirft2_sampled(pts, data[0], modes, -1, idx), out[0][idx], ref[idx]
    scale = np.abs(exact).max()
    e_ours, e_ref = np.abs(got - exact).max() / scale, np.abs(want - exact).max() / scale
    # Single precision with a low upsampling factor: the deconvolution divides by a phihat that spans ~1/tol, so the
    # fp32 rounding of the fine grid is amplified far above tol in BOTH implementations (measured: the reference
    # arithmetic itself is 1e-4 off the direct sum at sigma = 1.25, tol = 1e-6).  There the gate is "not worse than
    # the reference arithmetic", and parity is measured against that noise level instead of 1e-5.
    noisy = dtype == np.float32 and e_ref > 20 * tol
    assert rel_l2(out[0], ref) <= (max(TOL_PARITY[dtype], 3 * e_ref) if noisy else TOL_PARITY[dtype])
    # (the noise level itself moves with the order of the atomic adds from run to run: 1.9-2.2 x e_ref observed)
    assert e_ours <= max(20 * tol, 3e-6 if dtype == np.float32 else 1e-13, 3 * e_ref if noisy else 0.0)
    with pytest.raises(RuntimeError):
        gpu_nufft(nufft_type, modes, pts, data, tol, dtype, upsampfac=sigma, gpu_kerevalmeth=1)   # Horner needs sigma = 2
