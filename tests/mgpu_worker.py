"""One rank of the multi-GPU z-slab test (tests/test_mgpu_gpu.py): everything through the C ABI of
libcufinufft.so -- NCCL communicator, point routing, execute with its collectives -- with ctypes device
arrays (cufinufft_b200.gpuarray); no torch.  Usage: python mgpu_worker.py RANK WORLD RENDEZVOUS_FILE DTYPE
Rank r holds every WORLD-th point (offset r) of a seeded point set all ranks can regenerate; it checks its
share of the results against the undivided plan of the same library run on its own GPU and prints
'OK rank r ...' or raises."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(rank, world, path, dtype, modes=(24, 20, 32), M=60_000, tol=None):
    from cufinufft_b200 import cufinufft, gpuarray
    from cufinufft_b200.multi import MgpuComm, SlabPlan
    from helpers import cdtype, make_modes_data, make_points, make_strengths, rel_l2
    dtype = np.dtype(dtype).type
    tol = tol or (1e-5 if dtype == np.float32 else 1e-9)
    cd = cdtype(dtype)
    dev = rank % max(gpuarray.device_count(), 1)
    gpuarray.set_device(dev)
    comm = MgpuComm.file_rendezvous(world, rank, path, device=dev)
    shape = tuple(modes)[::-1]                     # python order (nZ, nY, nX); modes = (ms, mt, mu)
    pts = make_points(M, 3, dtype, seed=123, dist="wide")          # x, y, z over the whole valid range
    c_all = make_strengths(M, dtype, seed=5)[0]
    fk_all = make_modes_data(modes, dtype, seed=6)[0]
    mine = np.arange(rank, M, world)               # the points this rank HOLDS
    held = [np.ascontiguousarray(p[mine]) for p in pts]
    out = {}
    for nufft_type in (1, 2):
        # the undivided plan on this GPU: the truth for this library
        whole = cufinufft(nufft_type, shape, eps=tol, dtype=dtype, gpu_device_id=dev)
        dall = [gpuarray.to_gpu(p) for p in pts]
        whole.set_pts(*dall[::-1])
        if nufft_type == 1:
            fk_ref = gpuarray.zeros(shape, cd)
            whole.execute(gpuarray.to_gpu(c_all), fk_ref)
            want = fk_ref.get()
        else:
            c_ref = gpuarray.zeros((M,), cd)
            whole.execute(c_ref, gpuarray.to_gpu(fk_all))
            want = c_ref.get()[mine]
        whole.destroy()

        plan = SlabPlan(nufft_type, shape, eps=tol, dtype=dtype, rank=rank, world=world, gpu_device_id=dev)
        plan.set_comm(comm)
        dheld = [gpuarray.to_gpu(h) for h in held]
        n_owned = plan.route_set_pts(dheld[2], dheld[1], dheld[0])
        assert plan.info()["outside"] == 0
        if nufft_type == 1:
            c_held = gpuarray.to_gpu(np.ascontiguousarray(c_all[mine]))
            c_owned = gpuarray.zeros((max(n_owned, 1),), cd)
            plan.route_forward(c_held, c_owned)
            fk = gpuarray.zeros(shape, cd)
            for _ in range(2):                     # twice: the plan is reusable, buffers only grow
                plan.execute(c_owned, fk)
            got = fk.get()
        else:
            c_owned = gpuarray.zeros((max(n_owned, 1),), cd)
            plan.execute(c_owned, gpuarray.to_gpu(fk_all))
            c_back = gpuarray.zeros((max(mine.size, 1),), cd)
            plan.route_backward(c_owned, c_back)
            got = c_back.get()[:mine.size]
        err = rel_l2(got, want)
        out[nufft_type] = (err, n_owned)
        bound = 2e-6 if dtype == np.float32 else 1e-13
        assert err <= bound, (rank, nufft_type, err)
        plan.destroy()
    comm.destroy()
    print("OK rank %d/%d %s: type1 rel-l2 %.2e (owns %d), type2 rel-l2 %.2e (owns %d)"
          % (rank, world, np.dtype(dtype).name, out[1][0], out[1][1], out[2][0], out[2][1]), flush=True)


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4])
