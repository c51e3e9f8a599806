"""Multi-GPU partitioning of the hot path across one node (one process per GPU).

Two partitionings (BASELINE.json north_star):

* batched transforms (ntransf > 1): shard by TRANSFORM.  `c` is [ntransf][M] and `fk` is
  [ntransf][mu][mt][ms] with the transform index slowest (reference docs/cppdoc.md:165-167), so a
  contiguous block of transforms is a contiguous block of both arrays; every rank holds the same
  points, runs setpts itself and executes its block.  No collective on the data path.
* one large 3-D transform: z-slabs of the fine grid (`slab_range`, `slab_of_points`), points
  pre-binned by slab, halo planes exchanged between z-neighbours, one slab<->pencil all-to-all
  around the per-GPU FFTs.

This module is host logic only (pure Python/numpy, no CUDA): the device work goes through the
C ABI of libcufinufft.so; torch.distributed is used by the callers for the plumbing.
"""
import numpy as np


def transform_shard(ntransf, world, rank):
    """(first, count) of the transforms rank `rank` of `world` executes: contiguous blocks whose
    sizes differ by at most one, earlier ranks taking the larger ones."""
    if world < 1 or not 0 <= rank < world or ntransf < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(ntransf, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def slab_range(nf3, world, rank):
    """[z0, z1) planes of the fine grid owned by `rank`: contiguous, sizes differ by at most one."""
    first, count = transform_shard(nf3, world, rank)
    return first, first + count


def slab_halo(ns):
    """planes a slab needs from (type 2) / adds into (type 1) each z-neighbour: ceil(ns/2)."""
    return (ns + 1) // 2


def slab_of_points(z, nf3, world):
    """Owner rank of every point: the slab that contains floor(z_rescaled) with z_rescaled as
    RESCALE computes it (contrib/spreadinterp.h:36-38, double arithmetic narrowed to z.dtype)."""
    z = np.asarray(z)
    pi = z.dtype.type(np.pi)
    shift = np.where(z < -pi, 1.5, np.where(z >= pi, -0.5, 0.5))
    zr = ((z.astype(np.float64) * 0.159154943091895336 + shift) * nf3).astype(z.dtype)
    cell = np.clip(np.floor(zr).astype(np.int64), 0, nf3 - 1)
    bounds = np.array([slab_range(nf3, world, r)[1] for r in range(world)])
    return np.searchsorted(bounds, cell, side="right").astype(np.int32)
