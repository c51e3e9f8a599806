#!/usr/bin/env python
"""tools/compare_reference.py -- ours vs the reference library (cuFINUFFT v1.3 built for sm_100,
oracle/_ref/libcufinufft_ref.so) on the SAME device buffers: rel-l2 of the outputs and
exec-only / setpts timings (CUDA events via torch, median of `reps` after 2 warm-ups).
Prints one JSON line per config; results are recorded in BASELINE.md / profiles/.
  python tools/compare_reference.py [--configs 1,2,3,4,5] [--scale 1.0] [--reps 5]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import reflib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ref-method", type=int, default=0, help="override gpu_method for the reference (e.g. 4)")
    args = ap.parse_args()
    import torch
    from cufinufft_b200 import cufinufft
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    for ci in [int(c) for c in args.configs.split(",")]:
        cfg = dict(bench.CONFIGS[ci])
        M = int(cfg["M"] * args.scale)
        npdt = np.dtype(cfg["dtype"])
        tdt = torch.float32 if npdt == np.float32 else torch.float64
        npcd = np.complex64 if npdt == np.float32 else np.complex128
        shape = tuple(cfg["modes"])[::-1]
        nt = cfg["ntransf"]
        pts = bench.device_points(cfg, M, 42 + ci, torch, dev)
        if reflib.available():      # the reference reads an uninitialised weight for these
            from oracle import oracle as orc
            kp, nf, _, _ = orc.plan_params(cfg["type"], cfg["modes"], cfg["tol"], npdt, gpu_method=cfg["opts"].get("gpu_method"))
            pts = bench.drop_exact_stencil_points(pts, nf, kp.ns, torch)
            M = pts[0].numel()
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        c = torch.view_as_complex((torch.rand((nt, M, 2), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
        fk = torch.view_as_complex((torch.rand((nt,) + shape + (2,), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
        c2, fk2 = c.clone(), fk.clone()
        parr = [bench.TArr(p, npdt) for p in pts]
        out = {"config": cfg["name"], "M": M, "ntransf": nt}

        plan = cufinufft(cfg["type"], shape, n_trans=nt, eps=cfg["tol"], dtype=npdt, maxbatch=cfg.get("maxbatch", 1), **cfg["opts"])
        plan.set_stream(stream.cuda_stream)
        out["ours_setpts_ms"] = timed(lambda: plan.set_pts(*parr[::-1]), 3)
        ca, fa = bench.TArr(c, npcd), bench.TArr(fk, npcd)
        out["ours_exec_ms"] = timed(lambda: plan.execute(ca, fa), args.reps)
        plan.set_timing(True)
        plan.execute(ca, fa)
        out["ours_stages_ms"] = plan.timing()
        out["ours_pts_per_s"] = M * nt / (out["ours_exec_ms"] * 1e-3)
        plan.destroy()

        if reflib.available():
            ropts = dict(cfg["opts"])
            if args.ref_method:
                ropts["gpu_method"] = args.ref_method
            try:
                # the reference C API's own heuristic batch (maxbatchsize=0 -> min(ntransf, 8))
                ref = reflib.RefPlan(cfg["type"], cfg["modes"], cfg["tol"], npdt, ntransf=nt, maxbatch=0 if nt > 1 else 1, **ropts)
                out["ref_setpts_ms"] = timed(lambda: ref.set_pts(parr), 3)
                cb, fb = bench.TArr(c2, npcd), bench.TArr(fk2, npcd)
                out["ref_exec_ms"] = timed(lambda: ref.execute(cb, fb), args.reps)
                out["ref_pts_per_s"] = M * nt / (out["ref_exec_ms"] * 1e-3)
                out["speedup_exec"] = out["ref_exec_ms"] / out["ours_exec_ms"]
                out["speedup_setpts"] = out["ref_setpts_ms"] / out["ours_setpts_ms"]
                a, b = (fk, fk2) if cfg["type"] == 1 else (c, c2)
                out["rel_l2_vs_ref"] = float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())
                ref.destroy()
            except Exception as exc:  # noqa: BLE001
                out["ref_error"] = repr(exc)
        print(json.dumps(out), flush=True)
        del pts, c, fk, c2, fk2
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
