"""CPU tests of the z-slab decomposition of one 3-D transform: the
orchestration code of cufinufft_b200.multi (ring halo exchange, all-reduce of the partial mode
arrays, routing of points to their slabs) runs here on CPU tensors -- in one process with all
ranks emulated, and under torch.distributed (gloo, world_size 2 and 3) -- with the oracle-backed
slab stages of tests/slab_oracle.py standing in for the CUDA ones.  The result must equal the
undivided oracle transform."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import make_modes_data, make_points, make_strengths, rel_l2
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # modes (ms, mt, mu), M, tol, dtype, dist
    ((12, 10, 16), 3000, 1e-5, np.float32, "uniform"),
    ((10, 12, 14), 3000, 1e-9, np.float64, "wide"),
]


def _split(pts, data, nf3, world):
    from cufinufft_b200.multi import slab_of_points
    owner = slab_of_points(pts[2], nf3, world)
    return [np.flatnonzero(owner == r) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s" % ("x".join(map(str, c[0])), np.dtype(c[3]).name))
def test_emulated_slabs_type1_and_type2(case, world):
    import torch
    from cufinufft_b200.multi import slab_type1_emulated
    from slab_oracle import OracleSlab
    modes, M, tol, dtype, dist = case
    pts = make_points(M, 3, dtype, seed=21, dist=dist)
    shape = modes[::-1]
    tol_par = 2e-5 if dtype == np.float32 else 1e-12

    # type 1
    c = make_strengths(M, dtype)[0]
    plans = [OracleSlab(1, shape, tol, dtype, r, world) for r in range(world)]
    idx = _split(pts, c, plans[0].nf[2], world)
    assert sum(len(i) for i in idx) == M
    for p, i in zip(plans, idx):
        p.set_pts(pts[2][i], pts[1][i], pts[0][i])
    fk = torch.zeros(shape, dtype=torch.complex64 if dtype == np.float32 else torch.complex128)
    slab_type1_emulated(plans, [c[i] for i in idx], fk)
    want = orc.nufft(1, modes, pts, c, tol, dtype=dtype)
    assert rel_l2(fk.numpy(), want) <= tol_par

    # type 2
    fkin = make_modes_data(modes, dtype)[0]
    out = np.zeros(M, fkin.dtype)
    for r, i in enumerate(idx):
        p = OracleSlab(2, shape, tol, dtype, r, world)
        p.set_pts(pts[2][i], pts[1][i], pts[0][i])
        ci = np.zeros(len(i), fkin.dtype)
        p.type2(ci, fkin)
        out[i] = ci
    want = orc.nufft(2, modes, pts, fkin, tol, dtype=dtype)
    assert np.all(np.isfinite(out))
    assert rel_l2(out, want) <= tol_par


def test_slab_owner_numpy_and_torch_agree():
    import torch
    from cufinufft_b200.multi import slab_of_points, slab_of_points_torch
    rng = np.random.default_rng(3)
    for dt in (np.float32, np.float64):
        z = rng.uniform(-3 * np.pi, 3 * np.pi, 50000).astype(dt)
        for nf3, world in ((1024, 8), (30, 3), (512, 1)):
            a = slab_of_points(z, nf3, world)
            b = slab_of_points_torch(torch.from_numpy(z), nf3, world).numpy()
            assert np.array_equal(a, b)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from cufinufft_b200.multi import SlabRouter, slab_type1, slab_type2
    from slab_oracle import OracleSlab
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    modes, M, tol, dtype = (12, 10, 16), 4000, 1e-9, np.float64
    shape = modes[::-1]
    pts = make_points(M, 3, dtype, seed=33, dist="wide")
    c = make_strengths(M, dtype)[0]
    fkin = make_modes_data(modes, dtype)[0]
    held = np.arange(rank, M, world)                   # every rank HOLDS an arbitrary share of the points
    hp = [torch.from_numpy(p[held]) for p in pts]

    plan1 = OracleSlab(1, shape, tol, dtype, rank, world)
    router = SlabRouter(hp[2], plan1.nf[2], world, rank)
    own = router.forward(torch.stack(hp, dim=1))       # [n_owned][3] coordinates of the points this rank OWNS
    c_own = router.forward(torch.from_numpy(c[held]))
    plan1.set_pts(own[:, 2].contiguous(), own[:, 1].contiguous(), own[:, 0].contiguous())
    fk = torch.zeros(shape, dtype=torch.complex128)
    slab_type1(plan1, c_own, fk)

    plan2 = OracleSlab(2, shape, tol, dtype, rank, world)
    plan2.set_pts(own[:, 2].contiguous(), own[:, 1].contiguous(), own[:, 0].contiguous())
    c_out = torch.zeros(router.n_owned, dtype=torch.complex128)
    slab_type2(plan2, c_out, torch.from_numpy(fkin))
    back = router.backward(c_out)                      # values back at the holders, original order
    gathered = [None] * world
    dist.all_gather_object(gathered, (held, back.numpy()))
    if rank == 0:
        full = np.zeros(M, np.complex128)
        for h, v in gathered:
            full[h] = v
        np.save(os.path.join(out_dir, "t1.npy"), fk.numpy())
        np.save(os.path.join(out_dir, "t2.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])      # 2: both neighbours are the same peer; 3: the general ring
def test_slab_pipeline_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    modes, M, tol, dtype = (12, 10, 16), 4000, 1e-9, np.float64
    pts = make_points(M, 3, dtype, seed=33, dist="wide")
    c = make_strengths(M, dtype)[0]
    fkin = make_modes_data(modes, dtype)[0]
    t1 = np.load(os.path.join(str(tmp_path), "t1.npy"))
    t2 = np.load(os.path.join(str(tmp_path), "t2.npy"))
    assert rel_l2(t1, orc.nufft(1, modes, pts, c, tol, dtype=dtype)) <= 1e-12
    assert rel_l2(t2, orc.nufft(2, modes, pts, fkin, tol, dtype=dtype)) <= 1e-12
