#!/usr/bin/env python
"""tools/check_accuracy.py -- full-size accuracy audit on the GPU: ours and (when built) the
reference library against fp64 direct sums at sampled modes / targets (torch on the device is
only the checker here), NaN counts, and rel-l2 ours-vs-reference.  One JSON line per config.
  python tools/check_accuracy.py [--configs 1,2,3,4] [--scale 1.0] [--nsample 24]"""
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


def direct_type1(torch, pts, c, modes, iflag, sample):
    """fk[k] = sum_j c_j exp(i*iflag*k.x_j) at flat mode indices `sample` (x fastest)."""
    out = []
    c64 = c.to(torch.complex128)
    p64 = [p.to(torch.float64) for p in pts]
    for flat in sample:
        ks, rem = [], int(flat)
        for m in modes:
            ks.append(rem % m - m // 2)
            rem //= m
        phase = sum(k * p for k, p in zip(ks, p64))
        out.append(torch.sum(c64 * torch.exp(1j * iflag * phase)).item())
    return np.array(out)


def direct_type2(torch, pts, fk, modes, iflag, sample):
    """c_j = sum_k fk[k] exp(i*iflag*k.x_j) at point indices `sample`; fk is [mu][mt][ms]."""
    dev = fk.device
    f64 = fk.to(torch.complex128)
    out = []
    for j in sample:
        acc = f64
        for d, m in enumerate(modes):                    # contract x first (last axis)
            k = torch.arange(-(m // 2), (m - 1) // 2 + 1, device=dev, dtype=torch.float64)
            e = torch.exp(1j * iflag * k * float(pts[d][int(j)].item())).to(torch.complex128)
            acc = torch.tensordot(acc, e, dims=([acc.dim() - 1], [0]))
        out.append(acc.item())
    return np.array(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--nsample", type=int, default=24)
    args = ap.parse_args()
    import torch
    from cufinufft_b200 import cufinufft
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    for ci in [int(c) for c in args.configs.split(",")]:
        cfg = dict(bench.CONFIGS[ci])
        M = int(cfg["M"] * args.scale)
        npdt = np.dtype(cfg["dtype"])
        tdt = torch.float32 if npdt == np.float32 else torch.float64
        npcd = np.complex64 if npdt == np.float32 else np.complex128
        shape = tuple(cfg["modes"])[::-1]
        nt = cfg["ntransf"]
        iflag = 1 if cfg["type"] == 1 else -1
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
        parr = [bench.TArr(p, npdt) for p in pts]
        out = {"config": cfg["name"], "M": M, "ntransf": nt, "tol": cfg["tol"]}
        rng = np.random.default_rng(11)
        tsel = [0, nt - 1] if nt > 1 else [0]

        def audit(tag, cc, ff):
            res = cc if cfg["type"] == 2 else ff
            out[tag + "_nonfinite"] = int((~torch.isfinite(torch.view_as_real(res))).sum().item())
            errs = []
            for t in tsel:
                if cfg["type"] == 1:
                    sample = rng.integers(0, int(np.prod(cfg["modes"])), args.nsample)
                    exact = direct_type1(torch, pts, c[t], cfg["modes"], iflag, sample)
                    got = ff[t].reshape(-1)[torch.as_tensor(sample, device=dev)].cpu().numpy()
                else:
                    sample = rng.integers(0, M, args.nsample)
                    exact = direct_type2(torch, pts, fk[t], cfg["modes"], iflag, sample)
                    got = cc[t][torch.as_tensor(sample, device=dev)].cpu().numpy()
                errs.append((float(np.abs(got - exact).max() / np.abs(exact).max()),
                             float(np.linalg.norm(got - exact) / np.linalg.norm(exact))))
            out[tag + "_max_err_vs_direct"] = max(e[0] for e in errs)
            out[tag + "_rel_l2_vs_direct_sampled"] = max(e[1] for e in errs)

        plan = cufinufft(cfg["type"], shape, n_trans=nt, eps=cfg["tol"], dtype=npdt, maxbatch=cfg.get("maxbatch", 1), **cfg["opts"])
        plan.set_stream(stream.cuda_stream)
        plan.set_pts(*parr[::-1])
        c1, f1 = c.clone(), fk.clone()
        plan.execute(bench.TArr(c1, npcd), bench.TArr(f1, npcd))
        torch.cuda.synchronize()
        audit("ours", c1, f1)
        plan.destroy()
        if reflib.available():
            try:
                ref = reflib.RefPlan(cfg["type"], cfg["modes"], cfg["tol"], npdt, ntransf=nt, maxbatch=0 if nt > 1 else 1, **cfg["opts"])
                ref.set_pts(parr)
                c2, f2 = c.clone(), fk.clone()
                ref.execute(bench.TArr(c2, npcd), bench.TArr(f2, npcd))
                torch.cuda.synchronize()
                audit("ref", c2, f2)
                a, b = (f1, f2) if cfg["type"] == 1 else (c1, c2)
                out["rel_l2_ours_vs_ref"] = float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())
                per_t = torch.linalg.vector_norm((a - b).reshape(nt, -1), dim=1) / torch.linalg.vector_norm(b.reshape(nt, -1), dim=1)
                out["rel_l2_per_transform_max"] = float(per_t.max().item())
                ref.destroy()
            except Exception as exc:  # noqa: BLE001
                out["ref_error"] = repr(exc)
        print(json.dumps(out), flush=True)
        del pts, c, fk
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
