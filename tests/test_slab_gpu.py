"""GPU tests of the z-slab decomposition (csrc/slab.cu through the C ABI, cufinufft_b200.multi):
all ranks of a decomposition emulated on ONE device (world = 1, 2, 3: several slab plans, halo
buffers handed over directly) must reproduce the undivided plan of the same library (fp32 1e-5 /
fp64 1e-12 rel-l2, the north-star tolerances) and the CPU oracle; with >= 2 visible GPUs the same
pipeline runs as one process per GPU over NCCL (ring halo exchange, all-reduce, point routing)."""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import cdtype, gpu_nufft, make_modes_data, make_points, make_strengths, rel_l2
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL_PARITY = {np.float32: 1e-5, np.float64: 1e-12}

CASES = [
    # modes (ms, mt, mu), M, tol, dtype, dist, opts
    ((24, 20, 32), 30000, 1e-5, np.float32, "uniform", {}),                      # ns=6, SM spread / tile interp
    ((24, 20, 32), 30000, 1e-5, np.float32, "cluster", dict(gpu_method=1)),      # GM engines
    ((20, 18, 32), 20000, 1e-9, np.float64, "wide", {}),                         # config-5 shape: ns=10 fp64
    ((16, 12, 40), 5000, 1e-3, np.float32, "uniform", dict(gpu_sort=0, gpu_method=1)),
    # the coarse partition in front of the sort (csrc/setpts.cu; automatic only for key tables beyond L2, e.g. the
    # type-1 slabs of config 5's size) forced on slab plans: slab-local bins, global z in the records
    ((24, 20, 32), 30000, 1e-5, np.float32, "cluster", dict(_sort_levels=9)),
    ((20, 18, 32), 20000, 1e-9, np.float64, "wide", dict(_sort_levels=9)),
]


def _tensors(arrs, torch):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-%s" % ("x".join(map(str, c[0])), np.dtype(c[3]).name, c[4]))
def test_emulated_slabs_match_undivided_plan(case, world):
    import torch
    from cufinufft_b200.multi import SlabPlan, slab_of_points, slab_type1_emulated
    modes, M, tol, dtype, dist, opts = case
    shape = modes[::-1]
    cd = cdtype(dtype)
    tcd = torch.complex64 if dtype == np.float32 else torch.complex128
    pts = make_points(M, 3, dtype, seed=41, dist=dist)
    stream = torch.cuda.current_stream().cuda_stream

    for nufft_type in (1, 2):
        o = dict(opts)
        sort_levels = o.pop("_sort_levels", 0)
        if nufft_type == 2:
            o.pop("gpu_method", None)
        plans = [SlabPlan(nufft_type, shape, eps=tol, dtype=dtype, rank=r, world=world, **o) for r in range(world)]
        if sort_levels:
            for p in plans:
                p.set_sort_levels(sort_levels)
        g = plans[0].info()
        assert g["nf3"] >= 2 * modes[2] and g["pad"] == (g["ns"] + 1) // 2
        assert [p.info()["z0"] for p in plans] == [r * (g["nf3"] // world) + min(r, g["nf3"] % world) for r in range(world)]
        owner = slab_of_points(pts[2], g["nf3"], world)
        idx = [np.flatnonzero(owner == r) for r in range(world)]
        keep = []
        for p, i in zip(plans, idx):
            p.set_stream(stream)
            dev = _tensors([pts[2][i], pts[1][i], pts[0][i]], torch)
            p.set_pts(*dev)
            keep.append(dev)
            assert p.info()["outside"] == 0
        if nufft_type == 1:
            c = make_strengths(M, dtype)[0]
            fk = torch.zeros(shape, dtype=tcd, device="cuda")
            slab_type1_emulated(plans, _tensors([c[i] for i in idx], torch), fk)
            torch.cuda.synchronize()
            got = fk.cpu().numpy()
            whole = gpu_nufft(1, modes, pts, c[None], tol, dtype, **o)[0]
            ref = orc.nufft(1, modes, pts, c, tol, dtype=dtype)
        else:
            fkin = make_modes_data(modes, dtype)[0]
            fkd = torch.from_numpy(fkin).cuda()
            got = np.zeros(M, cd)
            for p, i in zip(plans, idx):
                ci = torch.zeros(max(len(i), 1), dtype=tcd, device="cuda")
                p.type2(ci, fkd)
                torch.cuda.synchronize()
                got[i] = ci.cpu().numpy()[: len(i)]
            whole = gpu_nufft(2, modes, pts, fkin[None], tol, dtype, **o)[0]
            ref = orc.nufft(2, modes, pts, fkin, tol, dtype=dtype)
        assert np.all(np.isfinite(got))
        assert rel_l2(got, whole) <= TOL_PARITY[dtype], (nufft_type, rel_l2(got, whole))
        assert rel_l2(got, ref) <= TOL_PARITY[dtype], (nufft_type, rel_l2(got, ref))
        for p in plans:
            p.destroy()


def test_points_outside_the_slab_are_counted():
    import torch
    from cufinufft_b200.multi import SlabPlan, slab_of_points
    dtype, modes = np.float32, (16, 16, 32)
    pts = make_points(4000, 3, dtype, seed=5)
    plan = SlabPlan(2, modes[::-1], eps=1e-4, dtype=dtype, rank=1, world=4)
    owner = slab_of_points(pts[2], plan.info()["nf3"], 4)
    plan.set_stream(torch.cuda.current_stream().cuda_stream)
    dev = _tensors([pts[2], pts[1], pts[0]], torch)
    plan.set_pts(*dev)                                     # ALL points, most of them foreign
    # foreign points whose z rounds onto the slab's upper edge plane are tolerated (stencil inside the halo)
    n_foreign = int(np.sum(owner != 1))
    assert 0.95 * n_foreign <= plan.info()["outside"] <= n_foreign
    counted = plan.info()["outside"]
    plan.set_sort_levels(9)                                # the same through the coarse partition: counted once, in its first pass
    plan.set_pts(*dev)
    assert plan.info()["outside"] == counted
    c = torch.zeros(4000, dtype=torch.complex64, device="cuda")
    fk = torch.from_numpy(make_modes_data(modes, dtype)[0]).cuda()
    plan.type2(c, fk)                                      # still in bounds: pulled onto the slab edge
    torch.cuda.synchronize()
    assert torch.isfinite(torch.view_as_real(c)).all()
    with pytest.raises(RuntimeError):
        SlabPlan(2, modes[::-1], eps=1e-4, dtype=dtype, rank=0, world=64)      # slabs thinner than the halo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from cufinufft_b200.multi import SlabPlan, SlabRouter, slab_type1, slab_type2
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    modes, M, tol, dtype = (24, 20, 32), 40000, 1e-9, np.float64
    shape = modes[::-1]
    pts = make_points(M, 3, dtype, seed=33, dist="wide")
    c = make_strengths(M, dtype)[0]
    fkin = make_modes_data(modes, dtype)[0]
    held = np.arange(rank, M, world)
    hp = [torch.from_numpy(p[held]).cuda() for p in pts]
    stream = torch.cuda.current_stream().cuda_stream

    plan1 = SlabPlan(1, shape, eps=tol, dtype=dtype, rank=rank, world=world, gpu_device_id=rank)
    plan1.set_stream(stream)
    router = SlabRouter(hp[2], plan1.info()["nf3"], world, rank)
    own = router.forward(torch.stack(hp, dim=1))
    xs = [own[:, k].contiguous() for k in range(3)]
    c_own = router.forward(torch.from_numpy(c[held]).cuda())
    plan1.set_pts(xs[2], xs[1], xs[0])
    fk = torch.zeros(shape, dtype=torch.complex128, device="cuda")
    slab_type1(plan1, c_own, fk)
    assert plan1.info()["outside"] == 0

    plan2 = SlabPlan(2, shape, eps=tol, dtype=dtype, rank=rank, world=world, gpu_device_id=rank)
    plan2.set_stream(stream)
    plan2.set_pts(xs[2], xs[1], xs[0])
    c_out = torch.zeros(max(router.n_owned, 1), dtype=torch.complex128, device="cuda")
    slab_type2(plan2, c_out, torch.from_numpy(fkin).cuda())
    back = router.backward(c_out[: router.n_owned])
    torch.cuda.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, (held, back.cpu().numpy()))
    if rank == 0:
        full = np.zeros(M, np.complex128)
        for h, v in gathered:
            full[h] = v
        np.save(os.path.join(out_dir, "t1.npy"), fk.cpu().numpy())
        np.save(os.path.join(out_dir, "t2.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_pipeline_nccl_multi_gpu(tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    modes, M, tol, dtype = (24, 20, 32), 40000, 1e-9, np.float64
    pts = make_points(M, 3, dtype, seed=33, dist="wide")
    c = make_strengths(M, dtype)[0]
    fkin = make_modes_data(modes, dtype)[0]
    t1 = np.load(os.path.join(str(tmp_path), "t1.npy"))
    t2 = np.load(os.path.join(str(tmp_path), "t2.npy"))
    assert rel_l2(t1, orc.nufft(1, modes, pts, c, tol, dtype=dtype)) <= 1e-12
    assert rel_l2(t2, orc.nufft(2, modes, pts, fkin, tol, dtype=dtype)) <= 1e-12
