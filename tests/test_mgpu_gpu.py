"""The multi-GPU z-slab path with its collectives INSIDE libcufinufft.so (csrc/mgpu.cu: NCCL communicator,
point routing, ring halo exchange, all-reduce), driven through the C ABI only -- no torch, one process per
GPU (tests/mgpu_worker.py).  World 1 runs on any GPU box (NCCL communicator of one rank: every code path except
the wire); worlds 2 and 4 run when that many GPUs are visible (`gpurun --gpus 2`).  Each rank holds every
WORLD-th point of the set, anywhere in the domain, and checks its results against the undivided plan of the
same library: <= 2e-6 (fp32) / 1e-13 (fp64) rel-l2 -- only the FFT factorisation differs."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def _ngpu():
    try:
        from cufinufft_b200 import gpuarray
        return gpuarray.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_slab_transform_with_library_side_collectives(world, dtype):
    if _ngpu() < world:
        pytest.skip("needs >= %d GPUs (gpurun --gpus %d)" % (world, world))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "nccl_id")
        procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), path, dtype], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=600)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(out)
        for r, (p, out) in enumerate(zip(procs, outs)):
            assert p.returncode == 0 and ("OK rank %d/%d" % (r, world)) in out, out[-3000:]


def test_library_links_nccl_and_not_torch():
    """`ldd libcufinufft.so`: NCCL is a dependency of the library itself now; torch and the oracle are not."""
    from cufinufft_b200 import _cufinufft as ll
    out = subprocess.run(["ldd", ll.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl" in out
    assert "torch" not in out and "oracle" not in out


def test_slab_owner_rule_matches_the_plan_ranges():
    """cell -> owner rank (mgpu.cu: slab_owner) is the inverse of multi.slab_range for every split."""
    from cufinufft_b200.multi import slab_range
    for nf3 in (16, 30, 100, 1024):
        for world in (1, 2, 3, 4, 7, 8):
            if nf3 // world < 1:
                continue
            base, extra = divmod(nf3, world)
            cut = extra * (base + 1)
            cell = np.arange(nf3)
            owner = np.where(cell < cut, cell // (base + 1), extra + (cell - cut) // max(base, 1))
            for r in range(world):
                z0, z1 = slab_range(nf3, world, r)
                assert np.all(owner[z0:z1] == r)
