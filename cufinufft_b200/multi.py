"""Multi-GPU partitioning of the hot path across one node (one process per GPU).

Two partitionings (BASELINE.json north_star):

* batched transforms (ntransf > 1): shard by TRANSFORM.  `c` is [ntransf][M] and `fk` is
  [ntransf][mu][mt][ms] with the transform index slowest (reference docs/cppdoc.md:165-167), so a
  contiguous block of transforms is a contiguous block of both arrays; every rank holds the same
  points, runs setpts itself and executes its block.  No collective on the data path.
* one large 3-D transform: z-slabs of the fine grid (csrc/slab.cu).  Rank r
  owns the planes `slab_range(nf3, world, r)` plus ceil(ns/2) halo planes per side and the points
  whose floor(z_rescaled) falls in its planes (`SlabRouter` moves points / strengths / values
  between the ranks that hold them and the ranks that own them).  The mode array is replicated:
    type 2  needs NO exchange (every rank derives its planes, halo included, from the modes);
    type 1  adds the spread halos into the z-neighbours (ring send/recv) and sums the per-rank
            partial mode arrays (all-reduce).
  `slab_type1` / `slab_type2` run one rank's part under torch.distributed (NCCL on the GPUs,
  gloo in the CPU tests); `slab_type1_emulated` runs all ranks of a decomposition in ONE process
  (several slab plans on one device, buffers handed over directly) for single-GPU tests.

The device work goes through the C ABI of libcufinufft.so (`SlabPlan`).  On GPUs the collectives are
INSIDE the library too (csrc/mgpu.cu: NCCL): `MgpuComm` + `SlabPlan.set_comm / route_set_pts /
route_forward / route_backward / execute` are ctypes calls and need no torch at all.  The torch.distributed
functions further down (`slab_type1`, `SlabRouter`, ...) are the round-1 orchestration kept for the gloo CPU
tests of the host logic and for callers that bring their own transport; torch is imported lazily there.
"""
import ctypes
from ctypes import byref, c_int, c_void_p

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


def slab_of_points_torch(z, nf3, world):
    """`slab_of_points` on a torch tensor (any device); int64 owner per point."""
    import torch
    pi = float(np.float32(np.pi)) if z.dtype == torch.float32 else float(np.pi)
    shift = torch.where(z < -pi, 1.5, torch.where(z >= pi, -0.5, 0.5)).to(torch.float64)
    zr = ((z.to(torch.float64) * 0.159154943091895336 + shift) * nf3).to(z.dtype)
    cell = torch.clamp(torch.floor(zr).to(torch.int64), 0, nf3 - 1)
    bounds = torch.tensor([slab_range(nf3, world, r)[1] for r in range(world)], device=z.device)
    return torch.searchsorted(bounds, cell, right=True)


def _rv(t):
    """real view of a complex torch tensor (gloo and NCCL both move real dtypes)"""
    import torch
    return torch.view_as_real(t) if t.is_complex() else t


def _ptr(a):
    """device (or host) address of a torch tensor / GPUArray-like object"""
    if a is None:
        return None
    return a.ptr if hasattr(a, "ptr") else a.data_ptr()


class MgpuComm:
    """NCCL communicator owned by the library (cufinufft_mgpu_comm_*), one per process.  `unique_id()` on rank 0
    gives the 128 bytes every rank passes to the constructor -- moving them between the processes is the
    caller's business (`file_rendezvous` below does it through a file; bench.py uses a torch.distributed
    broadcast)."""

    def __init__(self, world, rank, unique_id, device=0):
        from . import _cufinufft as _ll
        self._ll, self.world, self.rank, self.device = _ll, world, rank, device
        self.handle = c_void_p(None)
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        ier = _ll.mgpu_comm_create(world, rank, buf, device, byref(self.handle))
        if ier != 0:
            raise RuntimeError("Error creating the NCCL communicator (%d)." % ier)

    @staticmethod
    def unique_id():
        from . import _cufinufft as _ll
        buf = ctypes.create_string_buffer(128)
        if _ll.mgpu_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed.")
        return buf.raw

    @classmethod
    def file_rendezvous(cls, world, rank, path, device=0, timeout=120.0):
        """Rank 0 writes the unique id to `path` (atomically), the others wait for it."""
        import os
        import time
        if rank == 0:
            uid = cls.unique_id()
            with open(path + ".tmp", "wb") as fh:
                fh.write(uid)
            os.replace(path + ".tmp", path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > timeout:
                    raise RuntimeError("timed out waiting for " + path)
                time.sleep(0.05)
            uid = open(path, "rb").read()
        return cls(world, rank, uid, device)

    def destroy(self):
        if self.handle is not None and self.handle.value:
            self._ll.mgpu_comm_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class SlabPlan:
    """One rank's share of a z-slab-partitioned 3-D transform (C ABI: cufinufft*_slab_*).

    `modes` = (nZ, nY, nX) as for `cufinufft`; the slab axis is the FIRST (slowest) one, and it is
    the first array of `set_pts`.  Arrays are torch CUDA tensors or anything with `.ptr`."""

    def __init__(self, nufft_type, modes, eps=1e-6, isign=None, dtype=np.float64, rank=0, world=1, **kwargs):
        from . import _cufinufft as _ll
        self.plan = None
        if len(modes) != 3:
            raise ValueError("slab plans are 3-D")
        if isign is None:
            isign = -1 if nufft_type == 2 else +1
        self.dtype = np.dtype(dtype)
        if self.dtype == np.float64:
            self._fn, self.complex_dtype = _ll._api[""], np.complex128
        elif self.dtype == np.float32:
            self._fn, self.complex_dtype = _ll._api["f"], np.complex64
        else:
            raise TypeError("Expected np.float32 or np.float64.")
        self.type, self.rank, self.world = nufft_type, rank, world
        self.modes = tuple(int(m) for m in modes)
        opts = _ll.NufftOpts()
        if _ll._default_opts(nufft_type, 3, opts) != 0:
            raise RuntimeError("Configuration not yet implemented.")
        known = {name for name, _ in opts._fields_}
        for key, value in kwargs.items():
            if key not in known:
                raise TypeError(f"Invalid option '{key}'")
            setattr(opts, key, value)
        handle = c_void_p(None)
        cmodes = (c_int * 3)(*self.modes[::-1])
        ier = self._fn["slab_make_plan"](nufft_type, cmodes, isign, float(eps), rank, world, byref(handle), opts)
        if ier != 0:
            raise RuntimeError("Error creating slab plan (%d)." % ier)
        self.plan = handle
        self.references = []
        self.M = None
        self._halo = None

    # -- geometry ------------------------------------------------------------------------
    def info(self):
        v = (ctypes.c_longlong * 12)()
        if self._fn["slab_info"](self.plan, v) != 0:
            raise RuntimeError("Error reading slab geometry.")
        keys = ("z0", "z1", "pad", "nz_local", "nf1", "nf2", "nf3", "plane_cells", "rank", "world", "outside", "ns")
        return dict(zip(keys, [int(x) for x in v]))

    # -- stream / timing -----------------------------------------------------------------
    def set_stream(self, stream_handle):
        if self._fn["set_stream"](self.plan, c_void_p(int(stream_handle))) != 0:
            raise RuntimeError("Error setting stream.")

    def set_timing(self, on=True):
        self._fn["set_timing"](self.plan, int(bool(on)))

    def set_sort_levels(self, levels):
        """A/B and test switch of the point sort (cufinufft*_set_sort_levels; takes effect at the next set_pts)."""
        if self._fn["set_sort_levels"](self.plan, int(levels)) != 0:
            raise RuntimeError("Error setting the sort levels.")

    def timing(self):
        t = (ctypes.c_float * 5)()
        if self._fn["get_timing"](self.plan, t) != 0:
            raise RuntimeError("No timing recorded.")
        return dict(zip(("spread_interp_ms", "fft_ms", "deconv_amplify_ms", "memset_ms", "total_ms"), list(t)))

    def launch_counts(self):
        n = (c_int * 2)()
        self._fn["get_launch_counts"](self.plan, n)
        return dict(setpts=n[0], execute=n[1])

    # -- points --------------------------------------------------------------------------
    def set_pts(self, kz, ky, kx):
        """This rank's points only (slab axis first).  Kept referenced until the next call."""
        M = kz.numel() if hasattr(kz, "numel") else kz.size
        self.references = [kz, ky, kx]
        ier = self._fn["set_pts"](M, _ptr(kx), _ptr(ky), _ptr(kz), 0, None, None, None, self.plan)
        if ier != 0:
            raise RuntimeError("Error setting non-uniform points.")
        self.M = M

    # -- stages --------------------------------------------------------------------------
    def _ok(self, ier, what):
        if ier != 0:
            raise RuntimeError("Error in slab stage %s (%d)." % (what, ier))

    def type2(self, c, fk):
        self._ok(self._fn["slab_type2"](_ptr(c), _ptr(fk), self.plan), "type2")

    def type1_spread(self, c):
        self._ok(self._fn["slab_type1_spread"](_ptr(c), self.plan), "type1_spread")

    def halo_pack(self, side, buf):
        self._ok(self._fn["slab_halo_pack"](side, _ptr(buf), self.plan), "halo_pack")

    def halo_add(self, side, buf):
        self._ok(self._fn["slab_halo_add"](side, _ptr(buf), self.plan), "halo_add")

    def type1_finish(self, fk_partial):
        self._ok(self._fn["slab_type1_finish"](_ptr(fk_partial), self.plan), "type1_finish")

    # -- collectives inside the library (csrc/mgpu.cu) -----------------------------------------
    def set_comm(self, comm):
        self._comm = comm                                    # keep it alive as long as the plan
        self._ok(self._fn["slab_set_comm"](self.plan, comm.handle), "set_comm")

    def route_set_pts(self, kz, ky, kx):
        """The points this rank HOLDS (any z; slab axis first): routed to their owners and bin-sorted there.
        Returns the number of points this rank owns afterwards."""
        M = kz.numel() if hasattr(kz, "numel") else kz.size
        self.references = [kz, ky, kx]
        self._ok(self._fn["slab_route_setpts"](M, _ptr(kx), _ptr(ky), _ptr(kz), self.plan), "route_setpts")
        v = (ctypes.c_longlong * 2)()
        self._ok(self._fn["slab_route_info"](self.plan, v), "route_info")
        self.M = int(v[1])
        return self.M

    def route_forward(self, held, owned):
        self._ok(self._fn["slab_route_forward"](_ptr(held), _ptr(owned), self.plan), "route_forward")

    def route_backward(self, owned, held):
        self._ok(self._fn["slab_route_backward"](_ptr(owned), _ptr(held), self.plan), "route_backward")

    def execute(self, c, fk):
        """This rank's share of the transform, collectives included (type 1: `fk` is complete on return)."""
        self._ok(self._fn["slab_execute"](_ptr(c), _ptr(fk), self.plan), "execute")

    def halo_buffers(self, device=None):
        """(send_lo, send_hi, recv_prev, recv_next): four torch tensors of pad*nf1*nf2 complex numbers."""
        if self._halo is None:
            import torch
            g = self.info()
            n = g["pad"] * g["plane_cells"]
            cd = torch.complex64 if self.dtype == np.float32 else torch.complex128
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            self._halo = tuple(torch.empty(n, dtype=cd, device=dev) for _ in range(4))
        return self._halo

    def destroy(self):
        if self.plan is not None:
            self._fn["destroy_plan"](self.plan)
            self.plan = None
            self.references = []
            self._halo = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ---- orchestration: one rank under torch.distributed ---------------------------------------

def ring_exchange(send_lo, send_hi, recv_prev, recv_next, rank, world, group=None):
    """send_lo -> rank-1 (arrives there as recv_next), send_hi -> rank+1 (arrives as recv_prev);
    periodic.  world == 1: the two halos wrap onto the rank itself.  Tensors on any device."""
    if world == 1:
        recv_prev.copy_(send_hi)
        recv_next.copy_(send_lo)
        return
    import torch.distributed as dist
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    send_lo, send_hi, recv_prev, recv_next = _rv(send_lo), _rv(send_hi), _rv(recv_prev), _rv(recv_next)
    if dist.get_backend(group) == "nccl":
        # one grouped launch; with world == 2 both messages go to the same peer and are matched in
        # posting order (hi first, lo second on both sides)
        ops = [dist.P2POp(dist.isend, send_hi, nxt, group), dist.P2POp(dist.isend, send_lo, prev, group),
               dist.P2POp(dist.irecv, recv_prev, prev, group), dist.P2POp(dist.irecv, recv_next, nxt, group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return
    # gloo (CPU tests): tagged point-to-point
    reqs = [dist.isend(send_hi, nxt, group=group, tag=1), dist.isend(send_lo, prev, group=group, tag=2),
            dist.irecv(recv_prev, prev, group=group, tag=1), dist.irecv(recv_next, nxt, group=group, tag=2)]
    for r in reqs:
        r.wait()


def slab_type1(plan, c, fk, group=None):
    """Type 1 on one rank: `c` = strengths of this rank's points, `fk` = [mu][mt][ms] tensor that
    holds the COMPLETE mode array on return (summed over ranks).  `plan` is a SlabPlan or any object
    with the same stage methods (the CPU tests pass a numpy stand-in)."""
    plan.type1_spread(c)
    send_lo, send_hi, recv_prev, recv_next = plan.halo_buffers()
    plan.halo_pack(0, send_lo)
    plan.halo_pack(1, send_hi)
    ring_exchange(send_lo, send_hi, recv_prev, recv_next, plan.rank, plan.world, group)
    plan.halo_add(0, recv_prev)
    plan.halo_add(1, recv_next)
    plan.type1_finish(fk)
    if plan.world > 1:
        import torch.distributed as dist
        dist.all_reduce(_rv(fk), group=group)
    return fk


def slab_type2(plan, c, fk, group=None):
    """Type 2 on one rank: `fk` = the complete mode array (replicated), `c` receives the values at
    this rank's points.  No communication."""
    plan.type2(c, fk)
    return c


def slab_type1_emulated(plans, cs, fk):
    """All ranks of a decomposition in one process (plans[r] = rank r, cs[r] its strengths):
    buffers are handed from plan to plan directly; `fk` receives the sum.  Test helper."""
    world = len(plans)
    bufs = []
    for p, c in zip(plans, cs):
        p.type1_spread(c)
        b = p.halo_buffers()
        p.halo_pack(0, b[0])
        p.halo_pack(1, b[1])
        bufs.append(b)
    fk.zero_()
    part = fk.clone()
    for r, p in enumerate(plans):
        p.halo_add(0, bufs[(r - 1) % world][1])      # previous rank's high halo
        p.halo_add(1, bufs[(r + 1) % world][0])      # next rank's low halo
        p.type1_finish(part)
        fk += part
    return fk


# ---- routing points to their slabs ------------------------------------------------------

class SlabRouter:
    """Moves per-point data between the rank that HOLDS a point and the rank that OWNS it.

    Every rank passes the slab coordinate `z` of the points it holds.  `forward(t)` sends the rows
    of `t` (one row per held point: coordinates, strengths) to the owners and returns the rows this
    rank owns (grouped by source rank); `backward(t)` is the inverse (type-2 values back to the
    holders, original order).  Uses all_to_all_single on NCCL, tagged send/recv on gloo."""

    def __init__(self, z, nf3, world, rank, group=None):
        import torch
        self.world, self.rank, self.group = world, rank, group
        owner = slab_of_points_torch(z, nf3, world)
        self.order = torch.argsort(owner, stable=True)
        self.send_counts = torch.bincount(owner, minlength=world).to(torch.int64)
        if world == 1:
            self.recv_counts = self.send_counts.clone()
        else:
            import torch.distributed as dist
            gathered = [torch.empty_like(self.send_counts) for _ in range(world)]
            dist.all_gather(gathered, self.send_counts, group=group)
            self.recv_counts = torch.stack([g[rank] for g in gathered])
        self.n_held = int(z.numel())
        self.n_owned = int(self.recv_counts.sum().item())

    def _alltoallv(self, src, send_counts, recv_counts):
        import torch
        out = torch.empty((int(recv_counts.sum().item()),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        if self.world == 1:
            out.copy_(src)
            return out
        import torch.distributed as dist
        sc, rc = [int(v) for v in send_counts.tolist()], [int(v) for v in recv_counts.tolist()]
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all_single(_rv(out), _rv(src.contiguous()), rc, sc, group=self.group)
            return out
        so, ro = np.concatenate([[0], np.cumsum(sc)]), np.concatenate([[0], np.cumsum(rc)])
        reqs = []
        for peer in range(self.world):
            if peer == self.rank:
                out[ro[peer]:ro[peer + 1]] = src[so[peer]:so[peer + 1]]
                continue
            if sc[peer]:
                reqs.append(dist.isend(_rv(src[so[peer]:so[peer + 1]].contiguous()), peer, group=self.group, tag=7))
            if rc[peer]:
                reqs.append(dist.irecv(_rv(out[ro[peer]:ro[peer + 1]]), peer, group=self.group, tag=7))
        for r in reqs:
            r.wait()
        return out

    def forward(self, t):
        return self._alltoallv(t[self.order], self.send_counts, self.recv_counts)

    def backward(self, t):
        import torch
        back = self._alltoallv(t, self.recv_counts, self.send_counts)
        out = torch.empty_like(back)
        out[self.order] = back
        return out
