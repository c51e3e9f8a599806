"""CPU tests (no GPU): the oracle's spread / interp / deconvolve / bin-sort restatement checked
the way the reference's own drivers check the library -- against direct sums within the requested
tolerance (test/cufinufft2d1_test.cu:171-177, 2d2:182-189, 3d1:175-184, 3d2:191-200; python
tests/test_basic.py:31-40) -- plus structural properties (adjointness of spread/interp, bin-sort
invariants).  The GPU-side pin (oracle == reference library on the device) is tests/test_vs_reference_gpu.py."""
import numpy as np
import pytest

from helpers import make_modes_data, make_points, make_strengths
from oracle import oracle as orc

CASES = [
    (1, (40, 36), 3000, 1e-3, np.float32), (2, (40, 36), 3000, 1e-3, np.float32),
    (1, (40, 36), 3000, 1e-5, np.float32), (2, (40, 36), 3000, 1e-6, np.float32),
    (1, (40, 36), 3000, 1e-9, np.float64), (2, (40, 36), 3000, 1e-12, np.float64),
    (1, (16, 12, 10), 2000, 1e-5, np.float32), (2, (16, 12, 10), 2000, 1e-9, np.float64),
    (1, (120,), 1000, 1e-6, np.float32), (2, (120,), 1000, 1e-10, np.float64),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "t%d-%s-%g-%s" % (c[0], "x".join(map(str, c[1])), c[3], np.dtype(c[4]).name))
def test_oracle_transform_within_tol_of_direct_sum(case):
    t, modes, M, tol, dt = case
    pts = make_points(M, len(modes), dt, seed=3)
    rng = np.random.default_rng(9)
    if t == 1:
        c = make_strengths(M, dt)[0]
        fk = orc.nufft(1, modes, pts, c, tol, dtype=dt).ravel()
        idx = rng.integers(0, fk.size, 64)
        exact = orc.dirft1_sampled(pts, c, modes, 1, idx)
        got = fk[idx]
    else:
        fk = make_modes_data(modes, dt)[0]
        c = orc.nufft(2, modes, pts, fk, tol, dtype=dt)
        idx = rng.integers(0, M, 64)
        exact = orc.dirft2_sampled(pts, fk, modes, -1, idx)
        got = c[idx]
    floor = 3e-6 if dt == np.float32 else 1e-13
    assert np.abs(got - exact).max() / np.abs(exact).max() <= max(10 * tol, floor)


@pytest.mark.parametrize("horner", [0, 1])
def test_spread_and_interp_are_adjoint(horner):
    dt, modes, M = np.float64, (30, 28), 500
    pts = make_points(M, 2, dt, seed=4, dist="wide")          # full valid range [-3pi, 3pi)
    kp, nf, _, _ = orc.plan_params(1, modes, 1e-8, dt, kerevalmeth=horner)
    rng = np.random.default_rng(1)
    c = (rng.standard_normal(M) + 1j * rng.standard_normal(M))
    g = (rng.standard_normal(nf[::-1]) + 1j * rng.standard_normal(nf[::-1]))
    fw = orc.spread(pts, c, nf, kp)
    cg = orc.interp(pts, g, nf, kp)
    # <spread(c), g> == <c, interp(g)>  (real kernel)
    assert abs(np.vdot(g, fw) - np.vdot(cg, c)) <= 1e-12 * abs(np.vdot(g, fw))


def test_binsort_invariants():
    for dim, dt, dist in ((2, np.float32, "uniform"), (3, np.float32, "cluster"), (1, np.float64, "wide"), (3, np.float64, "onebin")):
        M = 5000
        pts = make_points(M, dim, dt, seed=6, dist=dist)
        nf = [96, 80, 64][:dim]
        bs = [[1024], [32, 32], [16, 16, 2]][dim - 1]
        out = orc.binsort(pts, nf, bs, maxsubprobsize=100)
        assert out["binsize"].sum() == M
        assert np.array_equal(out["binstartpts"], np.concatenate([[0], np.cumsum(out["binsize"])[:-1]]))
        assert np.array_equal(np.sort(out["idxnupts"]), np.arange(M))
        assert np.array_equal(out["numsubprob"], -(-out["binsize"] // 100))
        assert out["totalnumsubprob"] == out["numsubprob"].sum() == out["subprobstartpts"][-1]
        assert np.array_equal(np.repeat(np.arange(out["binsize"].size), out["numsubprob"]), out["subprob_to_bin"])


def test_empty_and_single_point_inputs():
    dt, modes = np.float64, (12, 10)
    kp, nf, _, _ = orc.plan_params(1, modes, 1e-6, dt)
    z = [np.zeros(0, dt), np.zeros(0, dt)]
    assert not orc.spread(z, np.zeros(0, np.complex128), nf, kp).any()
    fk = orc.nufft(1, modes, [np.array([0.3]), np.array([-1.1])], np.array([2.0 - 1.0j]), 1e-9, dtype=dt)
    k1 = np.arange(-6, 6)[None, :]
    k2 = np.arange(-5, 5)[:, None]
    exact = (2.0 - 1.0j) * np.exp(1j * (k1 * 0.3 + k2 * -1.1))
    assert np.abs(fk - exact).max() <= 1e-7          # tol 1e-9 relative to sum|c| ~ 2.2, edge modes
