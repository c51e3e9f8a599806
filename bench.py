#!/usr/bin/env python
"""bench.py -- NU points/s per execute of the cuFINUFFT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one execute() of the configured transform on synthetic inputs that are already
resident in HBM (`value`), or through the host-buffer C-ABI call with the H2D/D2H copies inside
the timed region (`e2e`).  Default workload = BASELINE.json configs[1] (2-D type 2 fp64,
2048^2 modes, M=4e7, tol 1e-9, GM-sort interp).  N>1: every rank runs one independent
transform of the same shape on its own points (a batch of N transforms sharded by
transform, no collective) -> weak scaling; config 4 (ntransf=64) is instead split by
transform across the ranks (strong).  Prints ONE JSON line on rank 0.

torch is used only for device buffers, streams/events and torch.distributed; all NUFFT work
goes through libcufinufft.so (C ABI).  --impl reference times the CPU oracle port (the
reference vendors no CPU spreader) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(name="cfg1: 2D type 1 fp32 1000x1000 M=1e7 uniform tol=1e-3 method 2 (SM)", type=1, modes=(1000, 1000),
            M=10_000_000, tol=1e-3, dtype="float32", dist="uniform", ntransf=1, opts=dict(gpu_method=2)),
    2: dict(name="cfg2: 2D type 2 fp64 2048x2048 M=4e7 uniform tol=1e-9 GM-sort interp", type=2, modes=(2048, 2048),
            M=40_000_000, tol=1e-9, dtype="float64", dist="uniform", ntransf=1, opts=dict(gpu_method=1, gpu_sort=1)),
    3: dict(name="cfg3: 3D type 1 fp32 256^3 M=1e8 clustered tol=1e-5 method 2 (SM)", type=1, modes=(256, 256, 256),
            M=100_000_000, tol=1e-5, dtype="float32", dist="blobs", ntransf=1, opts=dict(gpu_method=2)),
    4: dict(name="cfg4: 2D type 1 fp32 512x512 radial M=262144 ntransf=64 tol=1e-4", type=1, modes=(512, 512),
            M=262_144, tol=1e-4, dtype="float32", dist="radial", ntransf=64, opts=dict(gpu_method=2), maxbatch=0),
    5: dict(name="cfg5: 3D type 2 fp64 512^3 (1024^3 fine grid) M=1e9 uniform tol=1e-9, z-slab partitioned", type=2,
            modes=(512, 512, 512), M=1_000_000_000, tol=1e-9, dtype="float64", dist="uniform", ntransf=1,
            opts=dict(gpu_method=1, gpu_sort=1), slab=True),
    # config 4 names type 1 AND type 2: its type-2 half (tile interp, one launch per batch of 8)
    7: dict(name="cfg4-type2: 2D type 2 fp32 512x512 radial M=262144 ntransf=64 tol=1e-4", type=2, modes=(512, 512),
            M=262_144, tol=1e-4, dtype="float32", dist="radial", ntransf=64, opts=dict(gpu_method=1, gpu_sort=1), maxbatch=0),
    # not a BASELINE.json config: the type-1 twin of config 5, the path with the two collectives
    # (ring halo add + all-reduce of the mode array) -- run with --config 6
    6: dict(name="cfg5-type1: 3D type 1 fp64 512^3 (1024^3 fine grid) M=1e9 uniform tol=1e-9, z-slab partitioned", type=1,
            modes=(512, 512, 512), M=1_000_000_000, tol=1e-9, dtype="float64", dist="uniform", ntransf=1,
            opts=dict(gpu_method=2), slab=True),
}


SM_CLOCK_HZ = 1.965e9          # B200 boost clock the bench runs at (clocks.sm_mhz in the JSON line confirms it)


def ncu_traffic(config_id):
    """DRAM bytes of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[str(config_id)]["bytes"]
    except Exception:   # noqa: BLE001
        return None


def binding_roofline(cfg, stage, ns, M, nt, k_ms, num_sms=148):
    """The resource that actually bounds the dominant kernel: the HBM figure the
    contract asks for says little for these kernels, so the line also carries this one.
      interp (tile engine): every point reads its ns^d complex stencil values from shared memory:
        shared-memory read bandwidth, peak 128 B/clk/SM;
      spread: 2 ns^d FMAs per point on the FP32 (or FP64) pipe, peak 128 (64) FMA/clk/SM."""
    d = len(cfg["modes"])
    sF = 4 if cfg["dtype"] == "float32" else 8
    if stage == "interp":
        work = float(M) * nt * ns ** d * 2 * sF
        peak = 128.0 * num_sms * SM_CLOCK_HZ
        name, unit = "shared-memory read bandwidth", "GB/s"
    else:
        work = float(M) * nt * 2 * ns ** d
        peak = (128.0 if sF == 4 else 64.0) * num_sms * SM_CLOCK_HZ
        name, unit = "FP%d FMA pipe" % (8 * sF), "GFMA/s"
    achieved = work / (k_ms * 1e-3)
    return {"resource": name, "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": unit, "frac": achieved / peak,
            "peak_source": "nominal: per-SM rate x %d SMs x %.3f GHz" % (num_sms, SM_CLOCK_HZ / 1e9)}


def algorithmic_bytes(cfg, stage):
    """SURVEY.md 8(d) per-unit figures.  sF = sizeof(real), d = dim."""
    sF = 4 if cfg["dtype"] == "float32" else 8
    d = len(cfg["modes"])
    M, nt = cfg["M"], cfg["ntransf"]
    nmodes = int(np.prod(cfg["modes"]))
    nfcells = int(np.prod(cfg["nf"]))
    if stage in ("spread", "interp"):
        return nt * (M * (d * sF + 2 * sF + 4) + nfcells * 2 * sF)
    if stage == "deconvolve":
        return nt * 2 * nmodes * 2 * sF
    if stage == "amplify":
        return nt * (nmodes * 2 * sF + nfcells * 2 * sF)
    if stage == "setpts":
        return M * (2 * d * sF + 12)
    raise ValueError(stage)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def make_points_np(cfg, M, seed):
    rng = np.random.default_rng(seed)
    d = len(cfg["modes"])
    dt = np.dtype(cfg["dtype"])
    if cfg["dist"] == "uniform":
        return [rng.uniform(-np.pi, np.pi, M).astype(dt) for _ in range(d)]
    if cfg["dist"] == "blobs":
        cen = rng.uniform(-np.pi, np.pi, (8, d))
        which = rng.integers(0, 8, M)
        return [(np.mod(cen[which, k] + 0.05 * 2 * np.pi * rng.standard_normal(M) + np.pi, 2 * np.pi) - np.pi).astype(dt)
                for k in range(d)]
    if cfg["dist"] == "radial":
        nspoke = 512
        nsamp = M // nspoke
        r = np.linspace(-np.pi, np.pi, nsamp, endpoint=False)
        theta = np.arange(nspoke) * (np.pi * (3.0 - np.sqrt(5.0)))      # golden angle
        kx = np.outer(np.cos(theta), r).ravel()
        ky = np.outer(np.sin(theta), r).ravel()
        return [kx.astype(dt), ky.astype(dt)]
    raise ValueError(cfg["dist"])


def device_points(cfg, M, seed, torch, dev):
    """Synthetic coordinates generated ON the device (no host staging for 1e8 points)."""
    d = len(cfg["modes"])
    tdt = torch.float32 if cfg["dtype"] == "float32" else torch.float64
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if cfg["dist"] == "uniform":
        return [(torch.rand(M, generator=g, device=dev, dtype=tdt) * 2 - 1) * np.pi for _ in range(d)]
    if cfg["dist"] == "blobs":
        cen = (torch.rand((8, d), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
        which = torch.randint(0, 8, (M,), generator=g, device=dev)
        out = []
        for k in range(d):
            v = cen[which, k] + 0.05 * 2 * np.pi * torch.randn(M, generator=g, device=dev, dtype=torch.float64)
            out.append((torch.remainder(v + np.pi, 2 * np.pi) - np.pi).to(tdt).contiguous())
        return out
    pts = make_points_np(cfg, M, seed)
    return [torch.from_numpy(p).to(dev) for p in pts]


def drop_exact_stencil_points(pts, nf, ns, torch):
    """Device twin of tests/helpers.drop_exact_stencil_points: removes the points for which the
    REFERENCE reads an uninitialised kernel weight (x_r - ns/2 an exact integer) --
    its output is then garbage, often NaN.  Used only where ours is compared WITH the reference."""
    keep = torch.ones_like(pts[0], dtype=torch.bool)
    pi = torch.tensor(np.pi, dtype=pts[0].dtype).item()
    for d, x in enumerate(pts):
        shift = torch.where(x < -pi, 1.5, torch.where(x >= pi, -0.5, 0.5)).to(torch.float64)
        xr = ((x.to(torch.float64) * 0.159154943091895336 + shift) * nf[d]).to(x.dtype).to(torch.float64)
        t = xr - ns / 2.0
        keep &= torch.ceil(t) != t
    return [p[keep].contiguous() for p in pts]


class TArr:
    """torch tensor seen through the .ptr/.dtype/.size protocol of the Python binding."""

    def __init__(self, t, npdtype):
        self.t, self.ptr, self.dtype, self.size = t, t.data_ptr(), np.dtype(npdtype), t.numel()


def cpu_baseline(cfg, seconds_budget=15.0):
    """The oracle port on the host cores: fixed part (amplify|deconvolve + FFT) timed once,
    spread|interp timed on a bounded sample and extrapolated linearly in M."""
    from oracle import oracle as orc
    if cfg.get("slab") and int(np.prod(cfg["modes"])) > 128 ** 3:
        # config 5 on the host: the 1024^3 fine grid (17 GB) + numpy FFT does not fit the time box --
        # the point-proportional part is timed on a 128^3-mode problem and extrapolated in M
        cfg = dict(cfg, modes=(128, 128, 128))
    dt = np.dtype(cfg["dtype"])
    cd = np.complex64 if dt == np.float32 else np.complex128
    modes, dim = cfg["modes"], len(cfg["modes"])
    kp, nf, _, _ = orc.plan_params(cfg["type"], modes, cfg["tol"], dt)
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    Ms = min(cfg["M"], 200_000)
    pts = make_points_np(cfg, max(Ms, 512 * 8), 1)
    pts = [p[:Ms] for p in pts]
    fw = np.zeros(tuple(nf)[::-1], cd)
    # calibrate the sample size to ~seconds_budget
    t0 = time.perf_counter()
    if cfg["type"] == 1:
        orc.spread(pts, (rng.uniform(-1, 1, Ms) + 1j * rng.uniform(-1, 1, Ms)).astype(cd), nf, kp, fw)
    else:
        orc.interp(pts, fw, nf, kp)
    t_small = time.perf_counter() - t0
    Ms2 = int(min(cfg["M"], max(Ms, Ms * (0.6 * seconds_budget) / max(t_small, 1e-4))))
    pts = make_points_np(cfg, max(Ms2, 512 * 8), 2)
    pts = [p[:Ms2] for p in pts]
    t0 = time.perf_counter()
    if cfg["type"] == 1:
        orc.spread(pts, (rng.uniform(-1, 1, Ms2) + 1j * rng.uniform(-1, 1, Ms2)).astype(cd), nf, kp, fw)
    else:
        orc.interp(pts, fw, nf, kp)
    t_pts = time.perf_counter() - t0
    # fixed part
    kers = [orc.fwkerhalf(nf[d], kp) for d in range(dim)]
    fk = np.zeros(tuple(modes)[::-1], cd)
    t0 = time.perf_counter()
    orc.deconvolve(cfg["type"], fw, fk, modes, nf, kers, dt)
    fw2 = np.fft.fftn(fw)
    t_fixed = time.perf_counter() - t0
    del fw2
    t_full = t_fixed + t_pts * (cfg["M"] / Ms2)
    value = cfg["M"] * cfg["ntransf"] / (cfg["ntransf"] * t_full)
    return dict(value=value, unit="NU pts/s", cores=cores, kind="port",
                sample="oracle (C+OpenMP port of the reference arithmetic; the reference vendors no CPU spreader): "
                       "%s on %d of %d pts in %.2fs, + deconv/amplify+numpy FFT %.2fs once, extrapolated linearly in M"
                       % ("spread" if cfg["type"] == 1 else "interp", Ms2, cfg["M"], t_pts, t_fixed),
                seconds=t_pts + t_fixed + t_small)


def run_reference_impl(args, cfg, rank):
    if rank != 0:
        return
    res = None
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 2))):
        res = cpu_baseline(cfg, seconds_budget=12.0)
    line = {
        "impl": "reference", "metric": "NU points/s per execute", "value": res["value"], "unit": "NU pts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * cfg["M"] * cfg["ntransf"] / res["value"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if cfg["dtype"] == "float32" else "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "type": cfg["type"], "modes": list(cfg["modes"]), "M_per_gpu": cfg["M"],
                   "ntransf_per_gpu": cfg["ntransf"], "tol": cfg["tol"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "NU pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def run_slab(args, cfg, rank, local_rank, world):
    """Config 5: ONE 3-D type-2 transform, fine grid split into z-slabs over the ranks (strong
    scaling: M and the grid are fixed, every rank gets M/world points inside its slab and a
    replicated mode array).  Type 2 needs no collective (csrc/slab.cu)."""
    import torch
    import torch.distributed as dist
    from cufinufft_b200.multi import SlabPlan, slab_type1, slab_type2

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    npdt, tdt, cdt = np.dtype("float64"), torch.float64, torch.complex128
    shape = tuple(cfg["modes"])[::-1]
    M_total = cfg["M"]
    M = M_total // world + (1 if rank < M_total % world else 0)
    stream = torch.cuda.current_stream()
    ttype = cfg["type"]
    plan = SlabPlan(ttype, shape, eps=cfg["tol"], dtype=npdt, rank=rank, world=world, gpu_device_id=local_rank, **cfg["opts"])
    plan.set_stream(stream.cuda_stream)
    geo = plan.info()
    nf3, z0, z1 = geo["nf3"], geo["z0"], geo["z1"]
    g = torch.Generator(device=dev)
    g.manual_seed(4242 + rank)
    x = (torch.rand(M, generator=g, device=dev, dtype=tdt) * 2 - 1) * np.pi
    y = (torch.rand(M, generator=g, device=dev, dtype=tdt) * 2 - 1) * np.pi
    z = ((torch.rand(M, generator=g, device=dev, dtype=tdt) * ((z1 - z0) * (1 - 1e-12)) + z0) / nf3 - 0.5) * (2 * np.pi)
    gk = torch.Generator(device=dev)
    gk.manual_seed(7)                                   # the mode array is REPLICATED: same seed on every rank
    fk = torch.view_as_complex((torch.rand(shape + (2,), generator=gk, device=dev, dtype=tdt) * 2 - 1).contiguous())
    if ttype == 2:
        c = torch.zeros(M, dtype=cdt, device=dev)
    else:
        c = torch.view_as_complex((torch.rand((M, 2), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
        fk.zero_()

    def step(cc, ff):
        if ttype == 2:
            slab_type2(plan, cc, ff)
        else:
            slab_type1(plan, cc, ff)          # spread, ring halo add (NCCL), FFTs, all-reduce of the modes (NCCL)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    plan.set_pts(z, y, x)
    torch.cuda.synchronize()
    t_set = []
    for _ in range(3):
        ev[0].record(stream)
        plan.set_pts(z, y, x)
        ev[1].record(stream)
        torch.cuda.synchronize()
        t_set.append(ev[0].elapsed_time(ev[1]))
    setpts_ms = float(np.median(t_set))
    setpts_launches = plan.launch_counts()["setpts"]
    outside = plan.info()["outside"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    plan.set_timing(True)
    for _ in range(args.warmup):
        step(c, fk)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(c, fk)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = (plan.launch_counts()["execute"] + (1 if ttype == 1 else 0)) * args.steps
    stage_ms = []
    for _ in range(min(args.steps, 3)):
        step(c, fk)
        stage_ms.append(plan.timing())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    value = M_total / (ms_step * 1e-3)
    checksum = float(torch.view_as_real(c if ttype == 2 else fk).abs().sum().item())

    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        fk_host = torch.empty(shape, dtype=cdt, pin_memory=True)
        c_host = torch.empty(M, dtype=cdt, pin_memory=True)
        fk_host.copy_(fk)
        fk_dev = torch.empty_like(fk)
        torch.cuda.synchronize()

        c_dev = torch.empty_like(c)
        if ttype == 1:
            c_host.copy_(c)

        def e2e_step():
            if ttype == 2:
                fk_dev.copy_(fk_host, non_blocking=True)
                step(c, fk_dev)
                c_host.copy_(c, non_blocking=True)
            else:
                c_dev.copy_(c_host, non_blocking=True)
                step(c_dev, fk_dev)
                fk_host.copy_(fk_dev, non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        barrier()
        ksteps = 2
        e0.record(stream)
        for _ in range(ksteps):
            e2e_step()
        e1.record(stream)
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = float(te.item()) / ksteps
        nb_fk, nb_c = fk_host.numel() * 16, c_host.numel() * 16
        e2e = {"value": M_total / (ms_e2e * 1e-3), "unit": "NU pts/s", "h2d_bytes_per_step": nb_fk if ttype == 2 else nb_c,
               "d2h_bytes_per_step": nb_c if ttype == 2 else nb_fk, "ms_per_step": ms_e2e,
               "api": "per rank: pinned host input -> device, cufinufft_slab_* stages (C ABI), result -> pinned host"}
    except Exception as exc:   # noqa: BLE001
        e2e = {"value": None, "error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    k_ms = float(np.median([s["spread_interp_ms"] for s in stage_ms]))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    local_cells = geo["nz_local"] * geo["plane_cells"]
    abytes = M * (3 * 8 + 16 + 4) + local_cells * 16
    achieved = abytes / (k_ms * 1e-3) / 1e9
    stages = {k: float(np.median([s[k] for s in stage_ms])) for k in stage_ms[0]}
    line = {
        "metric": "NU points/s per execute", "value": value, "unit": "NU pts/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "type": ttype, "modes": list(cfg["modes"]), "M_total": M_total, "M_per_gpu": M,
                   "tol": cfg["tol"], "ns": geo["ns"], "fine_grid": [geo["nf1"], geo["nf2"], nf3],
                   "slab_planes_rank0": [z0, z1], "halo_planes": geo["pad"], "points_outside_slab": outside,
                   "l2": "inputs larger than L2 (no flush needed)",
                   "parallelism": "z-slab decomposition of the fine grid, points pre-binned by slab, mode array replicated; " +
                                  ("type 2: no collective (halo planes are derived locally)" if ttype == 2 else
                                   "type 1: ring halo add + all-reduce of the mode array over NCCL inside the timed step")},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "interp" if ttype == 2 else "spread", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": abytes, "kernel_ms": k_ms,
                     "binding": binding_roofline(cfg, "interp" if ttype == 2 else "spread", geo["ns"], M, 1, k_ms),
                     "note": "rank 0's interp launch; HBM roofline as the contract asks, `binding` is the resource that "
                             "bounds the kernel (DESIGN.md)"},
        "stages_ms": stages, "setpts": {"ms": setpts_ms, "pts_per_s": M / (setpts_ms * 1e-3), "launches": setpts_launches},
        "checksum_abs_c_rank0": checksum, "cpu_baseline": None,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(cfg)
        cb["sample"] = "fine grid down-scaled to 128^3 modes on the host: " + cb["sample"]
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--scale", type=float, default=1.0, help="scale M (debug)")
    ap.add_argument("--opt", action="append", default=[], help="override a cufinufft_opts field, e.g. --opt gpu_binsizex=8 (experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(CONFIGS[args.config])
    cfg["M"] = int(cfg["M"] * args.scale)
    if args.opt:
        cfg["opts"] = dict(cfg["opts"], **{k: int(v) for k, v in (o.split("=") for o in args.opt)})
        cfg["name"] += " [" + ",".join(args.opt) + "]"

    if args.impl == "reference":
        run_reference_impl(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    from cufinufft_b200 import cufinufft

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    if cfg.get("slab"):
        run_slab(args, cfg, rank, local_rank, world)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    npdt = np.dtype(cfg["dtype"])
    tdt = torch.float32 if npdt == np.float32 else torch.float64
    cdt = torch.complex64 if npdt == np.float32 else torch.complex128
    npcd = np.complex64 if npdt == np.float32 else np.complex128
    dim = len(cfg["modes"])
    M = cfg["M"]
    ntransf = cfg["ntransf"]
    strong = args.config in (4, 7) and world > 1
    if strong:
        ntransf = cfg["ntransf"] // world            # shard the batch by transform, no collective
    shape = tuple(cfg["modes"])[::-1]

    pts = device_points(cfg, M, 42 + args.config + 1000 * rank, torch, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    c = torch.view_as_complex((torch.rand((ntransf, M, 2), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
    fk = torch.view_as_complex((torch.rand((ntransf,) + shape + (2,), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())

    stream = torch.cuda.current_stream()
    opts = dict(cfg["opts"], gpu_device_id=local_rank)
    plan = cufinufft(cfg["type"], shape, n_trans=ntransf, eps=cfg["tol"], dtype=npdt, maxbatch=cfg.get("maxbatch", 1), **opts)
    plan.set_stream(stream.cuda_stream)
    geo = plan.geometry()
    cfg["nf"] = [geo["nf1"], geo["nf2"], geo["nf3"]][:dim]
    parr = [TArr(p, npdt) for p in pts]
    carr, fkarr = TArr(c, npcd), TArr(fk, npcd)

    # ---- setpts (reported separately) ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    plan.set_pts(*parr[::-1])
    torch.cuda.synchronize()
    t_set = []
    for _ in range(3):
        ev[0].record(stream)
        plan.set_pts(*parr[::-1])
        ev[1].record(stream)
        torch.cuda.synchronize()
        t_set.append(ev[0].elapsed_time(ev[1]))
    setpts_ms = float(np.median(t_set))
    setpts_launches = plan.launch_counts()["setpts"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps ----
    plan.set_timing(True)
    for _ in range(args.warmup):
        plan.execute(carr, fkarr)
    torch.cuda.synchronize()
    stage_ms = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        plan.execute(carr, fkarr)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = plan.launch_counts()["execute"] * args.steps
    # per-stage times (separate passes so the event reads do not perturb the timed loop)
    for _ in range(min(args.steps, 5)):
        plan.execute(carr, fkarr)
        stage_ms.append(plan.timing())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    units_per_step = M * ntransf * world              # whole-job NU points per step
    value = units_per_step / (ms_step * 1e-3)

    # ---- end-to-end: host (pinned) buffers through the host C-ABI call ----
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        c_host = torch.empty((ntransf, M), dtype=cdt, pin_memory=True)
        fk_host = torch.empty((ntransf,) + shape, dtype=cdt, pin_memory=True)
        c_host.copy_(c)
        fk_host.copy_(fk)
        torch.cuda.synchronize()
        fn = plan._fn["exec_host"]
        for _ in range(2):
            assert fn(c_host.data_ptr(), fk_host.data_ptr(), plan.plan) == 0
        barrier()
        ksteps = max(3, min(args.steps, 5))
        e0.record(stream)
        for _ in range(ksteps):
            assert fn(c_host.data_ptr(), fk_host.data_ptr(), plan.plan) == 0
        e1.record(stream)
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = float(te.item()) / ksteps
        nb_c, nb_fk = c_host.numel() * c_host.element_size(), fk_host.numel() * fk_host.element_size()
        e2e = {"value": units_per_step / (ms_e2e * 1e-3), "unit": "NU pts/s",
               "h2d_bytes_per_step": nb_c if cfg["type"] == 1 else nb_fk,
               "d2h_bytes_per_step": nb_fk if cfg["type"] == 1 else nb_c, "ms_per_step": ms_e2e,
               "api": "cufinufft[f]_execute_host (pinned host c/fk; H2D + execute + D2H per step)"}
    except Exception as exc:   # noqa: BLE001
        e2e = {"value": None, "error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (spread | interp), live CUDA-event duration ----
    stage = "spread" if cfg["type"] == 1 else "interp"
    k_ms = float(np.median([s["spread_interp_ms"] for s in stage_ms]))
    cfg_local = dict(cfg, ntransf=min(ntransf, geo["maxbatch"]))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    abytes = algorithmic_bytes(cfg_local, stage)
    achieved = abytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": stage, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(args.config) if args.scale == 1.0 and not args.opt else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes, "kernel_ms": k_ms,
                "binding": binding_roofline(cfg, stage, geo["ns"], M, min(ntransf, geo["maxbatch"]), k_ms),
                "note": "HBM roofline as the contract asks; `binding` is the resource that bounds this kernel (DESIGN.md)"}
    stages = {k: float(np.median([s[k] for s in stage_ms])) for k in stage_ms[0]}
    # the HBM-bound stage of the path (deconvolve | amplify): algorithmic bytes / live duration vs the measured copy peak
    dstage = "deconvolve" if cfg["type"] == 1 else "amplify"
    d_gbs = algorithmic_bytes(cfg_local, dstage) / (max(stages["deconv_amplify_ms"], 1e-6) * 1e-3) / 1e9
    stages_hbm = {dstage: {"achieved_gbs": d_gbs, "frac": d_gbs / peak}}

    line = {
        "metric": "NU points/s per execute", "value": value, "unit": "NU pts/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32" if npdt == np.float32 else "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "type": cfg["type"], "modes": list(cfg["modes"]), "M_per_gpu": M,
                   "ntransf_per_gpu": ntransf, "tol": cfg["tol"], "ns": geo["ns"], "fine_grid": cfg["nf"],
                   "gpu_method": geo["method"], "l2": "inputs larger than L2 (no flush needed)" if
                   (M * (dim + 2) * npdt.itemsize > 200e6) else "small working set: fits L2",
                   "parallelism": "one independent transform per rank (batch sharded by transform, no collective)"
                   if not strong else "ntransf sharded by transform across ranks, no collective"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "stages_ms": stages, "stages_hbm": stages_hbm,
        "setpts": {"ms": setpts_ms, "pts_per_s": M / (setpts_ms * 1e-3), "launches": setpts_launches,
                                        "hbm_frac": algorithmic_bytes(cfg, "setpts") / (setpts_ms * 1e-3) / 1e9 / peak},
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(cfg)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
