"""CPU test of the N>1 host logic with torch.distributed (gloo, world_size 2): the batch of
transforms is sharded by transform with no data-path collective, every rank computes its block
(the oracle stands in for the GPU here -- there is no device in this container), and the
gathered result must equal the single-process result bit for bit.  Also the slab partition."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from cufinufft_b200.multi import transform_shard
    from helpers import make_points, make_strengths
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    modes, M, ntransf, tol, dt = (24, 20), 600, 5, 1e-4, np.float32
    pts = make_points(M, 2, dt, seed=8)                 # every rank holds the same points
    c = make_strengths(M, dt, ntransf=ntransf)
    first, count = transform_shard(ntransf, world, rank)
    mine = np.stack([orc.nufft(1, modes, pts, c[t], tol, dtype=dt) for t in range(first, first + count)]) if count else \
        np.zeros((0,) + modes[::-1], np.complex64)
    # gather only to CHECK (the data path itself has no collective)
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, count, mine))
    if rank == 0:
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        np.save(os.path.join(out_dir, "sharded.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_transform_shard_partition():
    from cufinufft_b200.multi import transform_shard
    for ntransf in (0, 1, 5, 8, 64, 65):
        for world in (1, 2, 3, 4, 8):
            blocks = [transform_shard(ntransf, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == ntransf
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1
    with pytest.raises(ValueError):
        transform_shard(4, 2, 2)


def test_slab_partition_of_points():
    from cufinufft_b200.multi import slab_halo, slab_of_points, slab_range
    nf3, world = 1024, 8
    rng = np.random.default_rng(0)
    z = rng.uniform(-3 * np.pi, 3 * np.pi, 20000)
    owner = slab_of_points(z, nf3, world)
    assert owner.min() >= 0 and owner.max() < world
    zr = np.mod(z / (2 * np.pi) + 0.5, 1.0) * nf3
    for r in range(world):
        z0, z1 = slab_range(nf3, world, r)
        assert z1 - z0 == nf3 // world
        sel = owner == r
        assert np.all((np.floor(zr[sel]) >= z0 - 1e-9) & (np.floor(zr[sel]) < z1 + 1e-9))
    assert slab_halo(10) == 5 and slab_halo(7) == 4
    # every stencil plane of a point lies within its slab + halo
    ns = 10
    zs = np.ceil(zr - ns / 2.0)
    for r in range(world):
        z0, z1 = slab_range(nf3, world, r)
        sel = owner == r
        assert np.all(zs[sel] >= z0 - slab_halo(ns)) and np.all(zs[sel] + ns - 1 < z1 + slab_halo(ns))


def test_batch_sharded_by_transform_world2(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_points, make_strengths
    from oracle import oracle as orc
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npy"))
    modes, M, ntransf, tol, dt = (24, 20), 600, 5, 1e-4, np.float32
    pts = make_points(M, 2, dt, seed=8)
    c = make_strengths(M, dt, ntransf=ntransf)
    want = np.stack([orc.nufft(1, modes, pts, c[t], tol, dtype=dt) for t in range(ntransf)])
    assert got.shape == want.shape
    # same arithmetic on every rank -> identical up to the order of the oracle's OpenMP atomics
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 2e-6
