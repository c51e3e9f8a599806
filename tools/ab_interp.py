#!/usr/bin/env python
"""tools/ab_interp.py -- A/B of the two interpolation engines (gather vs shared-memory tile) on
type-2 problems shaped like the BASELINE.json configs: stage-only interp time (CUDA events,
median) and rel-l2 between the two results.
  python tools/ab_interp.py [case,case,...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CASES = [
    dict(name="cfg2 2D fp64 ns10 M4e7", modes=(2048, 2048), M=40_000_000, tol=1e-9, dtype="float64", dist="uniform"),
    dict(name="cfg1-shape type2 2D fp32 ns4 M1e7", modes=(1000, 1000), M=10_000_000, tol=1e-3, dtype="float32", dist="uniform"),
    dict(name="cfg4-shape type2 2D fp32 ns5 radial x64", modes=(512, 512), M=262_144, tol=1e-4, dtype="float32", dist="radial", ntransf=64),
    dict(name="cfg3-shape type2 3D fp32 ns6 M1e8 blobs", modes=(256, 256, 256), M=100_000_000, tol=1e-5, dtype="float32", dist="blobs"),
    dict(name="cfg5/16 3D fp64 ns10 256^3 M3e7", modes=(256, 256, 256), M=30_000_000, tol=1e-9, dtype="float64", dist="uniform"),
    # low densities (points per bin << block size): where does the tile engine stop paying?
    dict(name="3D fp64 ns10 256^3 M8e6 (30 pts/bin)", modes=(256, 256, 256), M=8_000_000, tol=1e-9, dtype="float64", dist="uniform"),
    dict(name="3D fp64 ns10 256^3 M2e6 (8 pts/bin)", modes=(256, 256, 256), M=2_000_000, tol=1e-9, dtype="float64", dist="uniform"),
    dict(name="3D fp32 ns6 256^3 M8e6 (30 pts/bin)", modes=(256, 256, 256), M=8_000_000, tol=1e-5, dtype="float32", dist="uniform"),
    dict(name="2D fp64 ns10 2048^2 M1e6 (61 pts/bin)", modes=(2048, 2048), M=1_000_000, tol=1e-9, dtype="float64", dist="uniform"),
    dict(name="2D fp32 ns4 1000^2 M2e5 (50 pts/bin)", modes=(1000, 1000), M=200_000, tol=1e-3, dtype="float32", dist="uniform"),
]


def main():
    import torch
    from cufinufft_b200 import cufinufft
    which = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else range(len(CASES))
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    for ci in which:
        cs = CASES[ci]
        npdt = np.dtype(cs["dtype"])
        tdt = torch.float32 if npdt == np.float32 else torch.float64
        npcd = np.complex64 if npdt == np.float32 else np.complex128
        nt = cs.get("ntransf", 1)
        shape = tuple(cs["modes"])[::-1]
        pts = bench.device_points(cs, cs["M"], 42 + ci, torch, dev)
        M = pts[0].numel()
        ntb = min(nt, 8)
        plan = cufinufft(2, shape, n_trans=nt, eps=cs["tol"], dtype=npdt, maxbatch=ntb, gpu_method=1)
        plan.set_stream(stream.cuda_stream)
        parr = [bench.TArr(p, npdt) for p in pts]
        plan.set_pts(*parr[::-1])
        geo = plan.geometry()
        nf = [geo["nf3"], geo["nf2"], geo["nf1"]][3 - len(shape):]
        g = torch.Generator(device=dev)
        g.manual_seed(3)
        fw = torch.view_as_complex((torch.rand((ntb,) + tuple(nf) + (2,), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
        outs, res = [], {"case": cs["name"], "M": M, "ns": geo["ns"]}
        for eng in (1, 2):
            plan.set_interp_engine(eng)
            c = torch.zeros((ntb, M), dtype=torch.complex64 if npdt == np.float32 else torch.complex128, device=dev)
            ca, fa = bench.TArr(c, npcd), bench.TArr(fw, npcd)
            for _ in range(2):
                plan.interp(ca, fa, ntb)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                plan.interp(ca, fa, ntb)
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res["engine%d_ms" % eng] = float(np.median(ts))
            res["engine%d_gpts" % eng] = M * ntb / np.median(ts) / 1e6
            outs.append(c)
        res["rel_l2_tile_vs_gather"] = float((torch.linalg.vector_norm(outs[1] - outs[0]) / torch.linalg.vector_norm(outs[0])).item())
        print(json.dumps(res), flush=True)
        plan.destroy()
        del pts, fw, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
