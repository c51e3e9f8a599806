"""Seeded random sweep over the plan space (type, dimension, precision, tolerance -> kernel width,
mode counts incl. odd and tiny ones, methods, sort flag, evaluator, bin sizes, maxsubprobsize,
batching, point distributions incl. the full [-3pi, 3pi) range): every draw must match the CPU oracle
to the parity tolerance and keep the reference-facing bin arrays bit-exact.  The fixed cases of
test_gpu_parity.py pin the BASELINE shapes; this sweep is for the corners between them (sub-bin
splits with odd bin sizes, grids smaller than a bin, one-point inputs, wide stencils)."""
import numpy as np
import pytest

from helpers import gpu_nufft, make_modes_data, make_points, make_strengths, rel_l2
from oracle import oracle as orc
from test_gpu_parity import TOL_PARITY, _check_bins

pytestmark = pytest.mark.gpu


def _draw(seed):
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.integers(1, 4))
    dtype = np.float32 if rng.random() < 0.5 else np.float64
    nufft_type = int(rng.integers(1, 3))
    tol = float(10.0 ** -rng.uniform(1.5, 6.0 if dtype == np.float32 else 13.0))
    top = {1: 400, 2: 70, 3: 26}[dim]
    modes = tuple(int(rng.integers(1, top)) for _ in range(dim))
    M = int(rng.choice([1, 2, 33, 1000, 20000, 120000 if dim < 3 else 60000]))
    dist = str(rng.choice(["uniform", "cluster", "wide", "onebin"]))
    opts = {}
    if rng.random() < 0.5:
        opts["gpu_method"] = int(rng.choice([1, 2]))
    if opts.get("gpu_method") == 1 and rng.random() < 0.3:
        opts["gpu_sort"] = 0
    if rng.random() < 0.3:
        opts["gpu_kerevalmeth"] = 1
    if rng.random() < 0.3:
        opts["gpu_maxsubprobsize"] = int(rng.choice([7, 100, 5000]))
    if rng.random() < 0.3:
        opts["gpu_binsizex"] = int(rng.choice([5, 8, 12, 24, 40]))
        if dim > 1:
            opts["gpu_binsizey"] = int(rng.choice([3, 8, 10, 32]))
        if dim > 2:
            opts["gpu_binsizez"] = int(rng.choice([1, 2, 3, 4]))
    ntransf = int(rng.choice([1, 1, 3]))
    maxbatch = int(rng.choice([0, 1, 2]))
    iflag = int(rng.choice([-1, 1]))
    return dict(dim=dim, dtype=dtype, type=nufft_type, tol=tol, modes=modes, M=M, dist=dist, opts=opts, ntransf=ntransf,
                maxbatch=maxbatch, iflag=iflag)


@pytest.mark.parametrize("seed", range(96))
def test_random_plan_matches_oracle(seed):
    d = _draw(seed)
    dim, dtype, modes, M = d["dim"], d["dtype"], d["modes"], d["M"]
    pts = make_points(M, dim, dtype, seed=500 + seed, dist=d["dist"])
    if d["type"] == 1:
        data = make_strengths(M, dtype, ntransf=d["ntransf"])
    else:
        data = make_modes_data(modes, dtype, ntransf=d["ntransf"])
    try:
        out, plan = gpu_nufft(d["type"], modes, pts, data, d["tol"], dtype, ntransf=d["ntransf"], maxbatch=d["maxbatch"],
                              iflag=d["iflag"], return_plan=True, **d["opts"])
    except RuntimeError:
        # a draw the library refuses (e.g. Horner with a width it has no table for) must be refused by
        # the oracle's parameter logic too -- the reference rejects it the same way
        kp = orc.KernelParams(d["tol"], dtype, 2.0, d["opts"].get("gpu_kerevalmeth", 0))
        assert kp.ier > 1, d
        return
    kem = d["opts"].get("gpu_kerevalmeth", 0)
    scale = 0.0
    for t in range(d["ntransf"]):
        ref = orc.nufft(d["type"], modes, pts, data[t], d["tol"], iflag=d["iflag"], dtype=dtype, kerevalmeth=kem)
        err = rel_l2(out[t], ref)
        scale = max(scale, float(np.abs(ref).max()))
        # one- and two-point inputs: the result can be tiny where the kernel tails cancel; rel-l2 is still meaningful.
        # all points in a few cells ("onebin"): the summation order of M additions per cell differs between the two
        # implementations (and between runs: atomics), noise floor ~ sqrt(M) * eps
        bound = TOL_PARITY[dtype]
        if d["dist"] == "onebin":
            bound = max(bound, 3 * np.sqrt(M) * np.finfo(dtype).eps)
        assert err <= bound, (d, t, err)
    assert np.all(np.isfinite(np.asarray(out).view(np.float32 if dtype == np.float32 else np.float64)))
    g = plan.geometry()
    if g["method"] == 2 or d["opts"].get("gpu_sort", 1):
        _check_bins(plan, pts, dtype, d["opts"].get("gpu_maxsubprobsize", 1024))
