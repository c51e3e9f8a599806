#!/usr/bin/env python
"""bench.py -- NU points/s per execute of the cuFINUFFT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..7] [--impl ours|reference] [--no-extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one execute() of the configured transform on synthetic inputs that are already resident
in HBM (`value`), or through the host-buffer C-ABI call with the H2D/D2H copies inside the timed
region (`e2e`).  Prints ONE JSON line on rank 0.

Default workloads (BASELINE.json):
  N = 1   config 3 -- 3-D type 1 fp32, 256^3 modes, M = 1e8 clustered points, tol 1e-5, SM spread: the
          north-star target config.  `extra` carries configs 1, 2, 4 (type 1 and 2) at full size and the
          one-GPU run of config 5 (the N = 1 point of its strong-scaling curve).
  N > 1   config 5 -- ONE 3-D type-2 fp64 transform, 512^3 modes, M = 1e9, z-slab partitioned over the N
          ranks (strong scaling).  `extra` carries config 4 with its 64 transforms sharded over the
          ranks (strong, no collective) and the type-1 twin of config 5 (ring halo add + all-reduce
          over NCCL inside the step).
Every single-GPU record also times the REFERENCE library (cuFINUFFT v1.3 built for sm_100 into
oracle/_ref/libcufinufft_ref.so) on the same device buffers: `vs_ref_gpu`.

torch is used only for device buffers, streams/events and torch.distributed; all NUFFT work goes
through libcufinufft.so (C ABI).  --impl reference times the CPU oracle port on the host cores (the
reference has no CPU implementation of this path and vendors no CPU spreader).
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
    # (ring halo add + all-reduce of the mode array)
    6: dict(name="cfg5-type1: 3D type 1 fp64 512^3 (1024^3 fine grid) M=1e9 uniform tol=1e-9, z-slab partitioned", type=1,
            modes=(512, 512, 512), M=1_000_000_000, tol=1e-9, dtype="float64", dist="uniform", ntransf=1,
            opts=dict(gpu_method=2), slab=True),
    # the adversarial distribution of the reference's spread tests: every point in one corner bin
    # (test/spread3d_test.cu:147-156), on config 3's shape
    8: dict(name="cfg3-onebin: 3D type 1 fp32 256^3 M=1e8 all points in one 16-cell corner (spread3d_test worst case) tol=1e-5 method 2",
            type=1, modes=(256, 256, 256), M=100_000_000, tol=1e-5, dtype="float32", dist="onebin", ntransf=1,
            opts=dict(gpu_method=2)),
    # profiling stand-in for config 5's kernel: same stencil, bins and point density (477 per bin) on an eighth of the grid
    9: dict(name="cfg5-eighth: 3D type 2 fp64 256^3 (512^3 fine grid) M=1.25e8 uniform tol=1e-9 (config 5's density, one GPU, undivided plan)",
            type=2, modes=(256, 256, 256), M=125_000_000, tol=1e-9, dtype="float64", dist="uniform", ntransf=1,
            opts=dict(gpu_method=1, gpu_sort=1)),
    # ... and of its type-1 twin (config 6): 3-D fp64 spreading with a 10-wide stencil at config 5's density
    10: dict(name="cfg6-eighth: 3D type 1 fp64 256^3 (512^3 fine grid) M=1.25e8 uniform tol=1e-9 method 2 (config 6's density, one GPU, undivided plan)",
             type=1, modes=(256, 256, 256), M=125_000_000, tol=1e-9, dtype="float64", dist="uniform", ntransf=1,
             opts=dict(gpu_method=2)),
}

METRIC = "NU points/s per execute"


# ---------------------------------------------------------------------------------------------
# measured peaks
# ---------------------------------------------------------------------------------------------
def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_SM_PEAKS = {}


def sm_peaks(device):
    """Shared-memory / FFMA2 / FFMA / DFMA peaks of this GPU from the library's own micro-benchmarks
    (csrc/microbench.cu), measured once per process."""
    if device not in _SM_PEAKS:
        from cufinufft_b200 import _cufinufft as ll
        names = {0: "smem_lds128_gbs", 1: "ffma2_gfma", 2: "ffma_gfma", 3: "dfma_gfma", 4: "smem_atomic_cas_gops", 5: "smem_lds64_gbs"}
        out = {}
        for what, name in names.items():
            v = ctypes.c_double(0.0)
            if ll.microbench(what, device, ctypes.byref(v)) == 0:
                out[name] = v.value / 1e9
        _SM_PEAKS[device] = out
    return _SM_PEAKS[device]


def ncu_profile(config_id):
    """Per-launch numbers of the dominant kernel from the committed ncu capture of this config
    (profiles/ncu_traffic.json): DRAM bytes and shared-memory wavefronts, or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[str(config_id)]
    except Exception:   # noqa: BLE001
        return {}


def binding_roofline(cfg, stage, ns, M, nt, k_ms, peaks, prof):
    """The SM resources that bound the dominant kernel, each against its MEASURED peak (csrc/microbench.cu):
      shared-memory pipe: wavefronts per launch from the committed ncu capture x 128 B / live duration;
      FMA pipe: 2 ns^d FMAs per point (spread: FFMA2 in fp32, DFMA in fp64; interp the same count).
    `frac` of the line is the larger of the two -- the resource that binds."""
    d = len(cfg["modes"])
    f32 = cfg["dtype"] == "float32"
    views = []
    fma_peak = peaks.get("ffma2_gfma" if f32 else "dfma_gfma")
    if fma_peak:
        ach = float(M) * nt * 2 * ns ** d / (k_ms * 1e-3) / 1e9
        views.append({"resource": "FP32 FMA pipe (FFMA2)" if f32 else "FP64 FMA pipe (DFMA)", "achieved": ach, "peak": fma_peak,
                      "unit": "GFMA/s", "frac": ach / fma_peak, "peak_source": "measured: cufinufft_b200_microbench"})
    wf = prof.get("smem_wavefronts")
    smem_peak = peaks.get("smem_lds128_gbs")
    if wf and smem_peak and prof.get("M") and abs(prof["M"] - M) <= 1e-3 * M:
        ach = wf * 128.0 / (k_ms * 1e-3) / 1e9
        views.append({"resource": "shared-memory pipe", "achieved": ach, "peak": smem_peak, "unit": "GB/s", "frac": ach / smem_peak,
                      "peak_source": "measured: cufinufft_b200_microbench (conflict-free LDS.128)",
                      "wavefronts_per_launch": wf, "wavefront_source": prof.get("source")})
    if not views:
        return None
    best = max(views, key=lambda v: v["frac"])
    return dict(best, views=views)


def algorithmic_bytes(cfg, stage, M=None, nt=None):
    """SURVEY.md 8(d) per-unit figures.  sF = sizeof(real), d = dim."""
    sF = 4 if cfg["dtype"] == "float32" else 8
    d = len(cfg["modes"])
    M = cfg["M"] if M is None else M
    nt = cfg["ntransf"] if nt is None else nt
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

    def finish(self):
        self.stop_flag = True
        self.join(timeout=3)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------
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
    if cfg["dist"] == "onebin":
        nf = 2 * cfg["modes"][0]
        return [(np.pi * rng.uniform(0, 1, M) / nf * 16).astype(dt) for _ in range(d)]
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
    if cfg["dist"] == "onebin":          # test/spread3d_test.cu:147-156: x = pi * rand01() / nf * 16
        nf = 2 * cfg["modes"][0]
        return [(torch.rand(M, generator=g, device=dev, dtype=tdt) * (np.pi / nf * 16)).contiguous() for _ in range(d)]
    pts = make_points_np(cfg, M, seed)
    return [torch.from_numpy(p).to(dev) for p in pts]


def drop_exact_stencil_points(pts, nf, ns, torch):
    """Device twin of tests/helpers.drop_exact_stencil_points: removes the points for which the
    REFERENCE reads an uninitialised kernel weight (x_r - ns/2 an exact integer) --
    its output is then garbage, often NaN.  Used wherever ours is compared WITH the reference."""
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


def config_record(cfg, M, ntransf, ns, nf, method, parallelism, extra=None):
    """The `config` object of a JSON line -- the same keys on our arm and on the reference arm."""
    dim = len(cfg["modes"])
    sF = 4 if cfg["dtype"] == "float32" else 8
    rec = {"workload": cfg["name"], "type": cfg["type"], "modes": list(cfg["modes"]), "M_per_gpu": M,
           "ntransf_per_gpu": ntransf, "tol": cfg["tol"], "ns": ns, "fine_grid": list(nf)[:dim], "gpu_method": method,
           "l2": "inputs larger than L2 (no flush needed)" if (M * ((dim + 2) * sF + 4 * sF) + 2 * sF * float(np.prod(list(nf)[:dim])) > 126e6)
                 else "small working set: fits L2",
           "parallelism": parallelism}
    if extra:
        rec.update(extra)
    return rec


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def _load_oracle():
    """The oracle with all host cores: torchrun exports OMP_NUM_THREADS=1, which would make the OpenMP
    port single-threaded -- set it before liboracle.so (libgomp) is loaded."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import oracle as orc
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(ctypes.c_int(cores))
    except OSError:
        pass
    return orc, cores


def cpu_baseline(cfg, seconds_budget=15.0):
    """One bounded sample of the workload on the host cores with the oracle port: the point-proportional
    stage (spread | interp) on Ms of the M points, the grid-proportional part (deconvolve | amplify + FFT)
    once on the full grid; the full-workload figure is the linear extrapolation in M (said so in `sample`)."""
    orc, cores = _load_oracle()
    full_modes = tuple(cfg["modes"])
    if cfg.get("slab") and int(np.prod(cfg["modes"])) > 128 ** 3:
        # config 5 on the host: the 1024^3 fine grid (17 GB) + numpy FFT does not fit the time box --
        # the point-proportional part is timed on a 128^3-mode problem and extrapolated in M
        cfg = dict(cfg, modes=(128, 128, 128))
    dt = np.dtype(cfg["dtype"])
    cd = np.complex64 if dt == np.float32 else np.complex128
    modes, dim = cfg["modes"], len(cfg["modes"])
    kp, nf, _, _ = orc.plan_params(cfg["type"], modes, cfg["tol"], dt)
    rng = np.random.default_rng(0)
    Ms = min(cfg["M"], 200_000)
    pts = make_points_np(cfg, max(Ms, 512 * 8), 1)
    pts = [p[:Ms] for p in pts]
    fw = np.zeros(tuple(nf)[::-1], cd)
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
    kers = [orc.fwkerhalf(nf[d], kp) for d in range(dim)]
    fk = np.zeros(tuple(modes)[::-1], cd)
    t0 = time.perf_counter()
    orc.deconvolve(cfg["type"], fw, fk, modes, nf, kers, dt)
    fw2 = np.fft.fftn(fw)
    t_fixed = time.perf_counter() - t0
    del fw2
    full = Ms2 >= cfg["M"] and tuple(modes) == full_modes
    t_full = t_fixed + t_pts * (cfg["M"] / Ms2)
    value = cfg["M"] * cfg["ntransf"] / (cfg["ntransf"] * t_full)
    stage = "spread" if cfg["type"] == 1 else "interp"
    sample = ("oracle port (C + OpenMP restatement of the reference arithmetic), %d threads: %s on %d of the %d points in %.2f s"
              " + deconvolve/amplify and numpy FFT of the %s grid once in %.2f s; %s"
              % (cores, stage, Ms2, cfg["M"], t_pts, "x".join(map(str, nf)), t_fixed,
                 "the whole workload was run" if full else "whole-workload figure extrapolated linearly in M"))
    return dict(value=value, unit="NU pts/s", cores=cores, kind="port", sample=sample, extrapolated=not full,
                seconds=t_pts + t_fixed + t_small, ns=kp.ns, nf=list(nf), sample_points=Ms2)


def run_reference_impl(args, cfg, cfg_id, rank):
    """--impl reference: the CPU oracle port on the host cores on our arm's config, metric and unit; each step a
    bounded sample (rank 0 only; the other ranks exit without work)."""
    if rank != 0:
        return
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    samples = [cpu_baseline(cfg, seconds_budget=10.0) for _ in range(steps)]
    res = max(samples, key=lambda s: s["value"])
    method = cfg["opts"].get("gpu_method", 2 if cfg["type"] == 1 else 1)
    line = {
        "impl": "reference", "reference_kind": "cpu_oracle_port", "metric": METRIC, "value": res["value"], "unit": "NU pts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_run": steps,
        "ms_per_step": 1e3 * cfg["M"] * cfg["ntransf"] / res["value"], "extrapolated": res["extrapolated"],
        "measured_s_per_sample": res["seconds"], "higher_is_better": True,
        "scaling": "strong" if cfg.get("slab") else "weak", "vs_baseline": None,
        "dtype": "f32" if cfg["dtype"] == "float32" else "f64", "data": "synthetic",
        "config": config_record(cfg, cfg["M"], cfg["ntransf"], res["ns"], res["nf"], method,
                                "host cores only (the reference GPU library is timed inside our arm: vs_ref_gpu)"),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")},
        "e2e": {"value": res["value"], "unit": "NU pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# the reference GPU library beside ours
# ---------------------------------------------------------------------------------------------
def time_reference_gpu(cfg, pts_arr, c, fk, ours_out, torch, stream, npdt, npcd, ours_exec_ms, ours_setpts_ms, reps=3):
    """cuFINUFFT v1.3 (sm_100 build, oracle/_ref) on the SAME device buffers: setpts and execute times (CUDA
    events, median of `reps` after one warm-up) and rel-l2 of its output against ours."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reflib
    if not reflib.available():
        return {"available": False, "why": "oracle/_ref/libcufinufft_ref.so not built"}
    nt = cfg["ntransf"] if "nt_local" not in cfg else cfg["nt_local"]
    try:
        # the reference C API's own batch heuristic (maxbatchsize = 0 -> min(ntransf, 8)); its default evaluator
        ref = reflib.RefPlan(cfg["type"], cfg["modes"], cfg["tol"], npdt, ntransf=nt, maxbatch=0 if nt > 1 else 1, **cfg["opts"])

        def timed(fn, n, stat=np.median):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(stat(ts))

        t_set = timed(lambda: ref.set_pts(pts_arr), 5)          # it cudaMallocs inside setpts: noisy, hence five calls
        t_set_min = timed(lambda: ref.set_pts(pts_arr), 3, np.min)
        c2 = c.clone() if cfg["type"] == 2 else c
        fk2 = fk.clone() if cfg["type"] == 1 else fk
        t_exec = timed(lambda: ref.execute(TArr(c2, npcd), TArr(fk2, npcd)), reps)
        a, b = (ours_out, fk2) if cfg["type"] == 1 else (ours_out, c2)
        rel = float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())
        ref.destroy()
        M = pts_arr[0].size
        return {"available": True, "library": "cuFINUFFT v1.3 built for sm_100 (oracle/_ref/libcufinufft_ref.so), same device buffers",
                "ref_exec_ms": t_exec, "ours_exec_ms": ours_exec_ms, "speedup_exec": t_exec / ours_exec_ms,
                "ref_setpts_ms": t_set, "ref_setpts_min_ms": min(t_set, t_set_min), "ours_setpts_ms": ours_setpts_ms, "speedup_setpts": t_set / ours_setpts_ms,
                "ref_pts_per_s": M * nt / (t_exec * 1e-3), "rel_l2_ours_vs_ref": rel,
                "evaluator": "gpu_kerevalmeth=%d on both sides" % cfg["opts"].get("gpu_kerevalmeth", 0)}
    except Exception as exc:   # noqa: BLE001
        return {"available": False, "why": repr(exc)}


# ---------------------------------------------------------------------------------------------
# one transform per rank / a batch sharded by transform
# ---------------------------------------------------------------------------------------------
def run_plain(args, cfg_id, cfg, ctx, steps, warmup, with_e2e=True, with_cpu=True, with_ref=True):
    torch, dist = ctx["torch"], ctx["dist"]
    from cufinufft_b200 import cufinufft
    rank, local_rank, world, dev = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dev"]
    cfg = dict(cfg)
    npdt = np.dtype(cfg["dtype"])
    tdt = torch.float32 if npdt == np.float32 else torch.float64
    cdt = torch.complex64 if npdt == np.float32 else torch.complex128
    npcd = np.complex64 if npdt == np.float32 else np.complex128
    dim = len(cfg["modes"])
    ntransf = cfg["ntransf"]
    strong = cfg["ntransf"] > 1 and world > 1
    if strong:
        ntransf = cfg["ntransf"] // world            # shard the batch by transform, no collective
    shape = tuple(cfg["modes"])[::-1]
    stream = torch.cuda.current_stream()
    opts = dict(cfg["opts"], gpu_device_id=local_rank)
    plan = cufinufft(cfg["type"], shape, n_trans=ntransf, eps=cfg["tol"], dtype=npdt, maxbatch=cfg.get("maxbatch", 1), **opts)
    plan.set_stream(stream.cuda_stream)
    if args.sort_levels:
        plan.set_sort_levels(args.sort_levels)
    geo = plan.geometry()
    cfg["nf"] = [geo["nf1"], geo["nf2"], geo["nf3"]][:dim]

    pts = device_points(cfg, cfg["M"], 42 + cfg_id + (0 if strong else 1000 * rank), torch, dev)
    # the few points for which the reference reads an uninitialised kernel weight are removed, so that both
    # libraries can be timed and compared on identical inputs (~1e-4 of fp32 inputs, none in fp64)
    pts = drop_exact_stencil_points(pts, cfg["nf"], geo["ns"], torch)
    M = pts[0].numel()
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    c = torch.view_as_complex((torch.rand((ntransf, M, 2), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
    fk = torch.view_as_complex((torch.rand((ntransf,) + shape + (2,), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
    parr = [TArr(p, npdt) for p in pts]
    carr, fkarr = TArr(c, npcd), TArr(fk, npcd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    # ---- device-resident steps ----
    plan.set_timing(True)
    for _ in range(warmup):
        plan.execute(carr, fkarr)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        plan.execute(carr, fkarr)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = plan.launch_counts()["execute"] * steps
    stage_ms = []
    for _ in range(min(steps, 5)):                     # per-stage times in separate passes (event reads do not perturb the timed loop)
        plan.execute(carr, fkarr)
        stage_ms.append(plan.timing())
    clocks = sampler.finish()
    tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / steps
    units_per_step = M * ntransf * world              # whole-job NU points per step
    value = units_per_step / (ms_step * 1e-3)

    # ---- end-to-end: host (pinned) buffers through the host C-ABI call ----
    e2e = None
    if with_e2e and not args.no_e2e:
        try:
            c_host = torch.empty((ntransf, M), dtype=cdt, pin_memory=True)
            fk_host = torch.empty((ntransf,) + shape, dtype=cdt, pin_memory=True)
            c_host.copy_(c)
            fk_host.copy_(fk)
            torch.cuda.synchronize()
            # a host-buffer caller also hands the points over from the host (cufinufft*_setpts_host: besides the
            # plan's own sort the points are sorted in chunks of the caller's index range, so that execute_host
            # can overlap the PCIe copies with the kernels chunk by chunk); setpts is not part of the timed step
            pts_host = [p_.cpu().numpy() for p_ in pts]
            plan.set_pts_host(*pts_host[::-1])
            torch.cuda.synchronize()
            fn = plan._fn["exec_host"]
            for _ in range(2):
                assert fn(c_host.data_ptr(), fk_host.data_ptr(), plan.plan) == 0
            barrier()
            ksteps = max(3, min(steps, 5))
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
                   "api": "cufinufft[f]_setpts_host once, then cufinufft[f]_execute_host per step (pinned host c/fk; H2D, kernels and D2H inside the timed call, pipelined in chunks)"}
            del c_host, fk_host, pts_host
        except Exception as exc:   # noqa: BLE001
            e2e = {"value": None, "error": repr(exc)}

    # ---- the reference library on the same buffers (rank 0 only; single-transform-per-rank workloads) ----
    vs_ref = None
    if with_ref and rank == 0 and not args.no_ref:
        plan.execute(carr, fkarr)
        torch.cuda.synchronize()
        ours_out = (fk if cfg["type"] == 1 else c).clone()
        exec_ms = float(np.median([s["total_ms"] for s in stage_ms])) if ntransf <= geo["maxbatch"] else ms_step
        vs_ref = time_reference_gpu(dict(cfg, nt_local=ntransf), parr, c, fk, ours_out, torch, stream, npdt, npcd, exec_ms, setpts_ms,
                                    reps=1 if cfg.get("dist") == "onebin" else 3)
        del ours_out
    barrier()

    line = None
    if rank == 0:
        stage = "spread" if cfg["type"] == 1 else "interp"
        k_ms = float(np.median([s["spread_interp_ms"] for s in stage_ms]))
        nt_launch = min(ntransf, geo["maxbatch"])
        peak, peak_src = hbm_peak()
        abytes = algorithmic_bytes(cfg, stage, M, nt_launch)
        achieved = abytes / (k_ms * 1e-3) / 1e9
        prof = ncu_profile(cfg_id)
        roofline = {"bound": "hbm", "kernel": stage, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": prof.get("bytes") if prof.get("M") and abs(prof["M"] - M) <= 1e-3 * M else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes, "kernel_ms": k_ms,
                    "binding": binding_roofline(cfg, stage, geo["ns"], M, nt_launch, k_ms, sm_peaks(local_rank), prof),
                    "note": "HBM roofline as the contract asks; `binding` = the SM resource that bounds this kernel, against its measured peak"}
        stages = {k: float(np.median([s[k] for s in stage_ms])) for k in stage_ms[0]}
        dstage = "deconvolve" if cfg["type"] == 1 else "amplify"
        d_gbs = algorithmic_bytes(cfg, dstage, M, nt_launch) / (max(stages["deconv_amplify_ms"], 1e-6) * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "NU pts/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32" if npdt == np.float32 else "f64", "data": "synthetic",
            "config": config_record(cfg, M, ntransf, geo["ns"], cfg["nf"], geo["method"],
                                    "ntransf sharded by transform across ranks, no collective" if strong else
                                    "one independent transform per rank (replicas, no collective)"),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "vs_ref_gpu": vs_ref,
            "stages_ms": stages, "stages_hbm": {dstage: {"achieved_gbs": d_gbs, "frac": d_gbs / peak}},
            "setpts": {"ms": setpts_ms, "pts_per_s": M / (setpts_ms * 1e-3), "launches": setpts_launches,
                       "hbm_frac": algorithmic_bytes(cfg, "setpts", M) / (setpts_ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": None,
        }
        if with_cpu and not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(cfg)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")}
    plan.destroy()
    del pts, c, fk
    torch.cuda.empty_cache()
    return line


# ---------------------------------------------------------------------------------------------
# ONE 3-D transform, fine grid split into z-slabs over the ranks
# ---------------------------------------------------------------------------------------------
def run_slab(args, cfg_id, cfg, ctx, steps, warmup, with_e2e=True, with_cpu=True):
    """Strong scaling: M and the grid are fixed; every rank HOLDS M/world points anywhere in the domain, the
    timed `setpts` includes routing them to the ranks that own their slabs (SlabRouter: counts all-gather +
    all-to-all) and the bin sort there.  Type 2 then needs no collective; type 1 adds the halos around the
    ring and all-reduces the mode array inside the step."""
    torch, dist = ctx["torch"], ctx["dist"]
    from cufinufft_b200.multi import MgpuComm, SlabPlan
    rank, local_rank, world, dev = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["dev"]
    npdt, tdt, cdt = np.dtype("float64"), torch.float64, torch.complex128
    shape = tuple(cfg["modes"])[::-1]
    M_total = cfg["M"]
    M_held = M_total // world + (1 if rank < M_total % world else 0)
    stream = torch.cuda.current_stream()
    ttype = cfg["type"]
    plan = SlabPlan(ttype, shape, eps=cfg["tol"], dtype=npdt, rank=rank, world=world, gpu_device_id=local_rank, **cfg["opts"])
    plan.set_stream(stream.cuda_stream)
    geo = plan.info()
    nf3 = geo["nf3"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    g = torch.Generator(device=dev)
    g.manual_seed(4242 + rank)
    held = [(torch.rand(M_held, generator=g, device=dev, dtype=tdt) * 2 - 1) * np.pi for _ in range(3)]   # z, y, x of the points this rank holds

    # the library's own NCCL communicator: only the 128-byte id travels through torch.distributed (plumbing)
    uid = [MgpuComm.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = MgpuComm(world, rank, uid[0], device=local_rank)
    plan.set_comm(comm)

    # ---- setpts = route the held points to their owners + bin sort, all inside the library (timed together) ----
    plan.route_set_pts(held[0], held[1], held[2])
    torch.cuda.synchronize()
    t_set, t_sort = [], []
    for _ in range(3):
        barrier()
        e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        e0.record(stream)
        plan.route_set_pts(held[0], held[1], held[2])
        e1.record(stream)
        torch.cuda.synchronize()
        t_set.append(e0.elapsed_time(e1))
    M = plan.M
    t_sort = [0.0]
    ts = torch.tensor([float(np.median(t_set)), float(np.median(t_sort))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    setpts_ms, sort_ms = float(ts[0].item()), float(ts[1].item())
    setpts_launches = plan.launch_counts()["setpts"] + 9       # + owner, slot, 3 x pack kernels and the NCCL calls of routing
    outside = plan.info()["outside"]

    gk = torch.Generator(device=dev)
    gk.manual_seed(7)                                   # the mode array is REPLICATED: same seed on every rank
    fk = torch.view_as_complex((torch.rand(shape + (2,), generator=gk, device=dev, dtype=tdt) * 2 - 1).contiguous())
    if ttype == 2:
        c = torch.zeros(M, dtype=cdt, device=dev)
    else:
        c = torch.view_as_complex((torch.rand((M, 2), generator=g, device=dev, dtype=tdt) * 2 - 1).contiguous())
        fk.zero_()

    def step(cc, ff):
        plan.execute(cc, ff)                  # type 1: spread, ring halo add (NCCL), FFTs, all-reduce of the modes (NCCL) -- one C call

    plan.set_timing(True)
    for _ in range(warmup):
        step(c, fk)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step(c, fk)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = (plan.launch_counts()["execute"] + (1 if ttype == 1 else 0)) * steps
    stage_ms = []
    for _ in range(min(steps, 3)):
        step(c, fk)
        stage_ms.append(plan.timing())
    clocks = sampler.finish()
    tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / steps
    value = M_total / (ms_step * 1e-3)
    checksum = float(torch.view_as_real(c if ttype == 2 else fk).abs().sum().item())

    e2e = None
    if with_e2e and not args.no_e2e:
        try:
            fk_host = torch.empty(shape, dtype=cdt, pin_memory=True)
            c_host = torch.empty(M, dtype=cdt, pin_memory=True)
            fk_host.copy_(fk)
            fk_dev = torch.empty_like(fk)
            c_dev = torch.empty_like(c)
            if ttype == 1:
                c_host.copy_(c)
            torch.cuda.synchronize()

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
                   "api": "per rank: pinned host input -> device, cufinufft_slab_execute (C ABI, NCCL inside), result -> pinned host"}
            del fk_host, c_host, fk_dev, c_dev
        except Exception as exc:   # noqa: BLE001
            e2e = {"value": None, "error": repr(exc)}

    line = None
    if rank == 0:
        k_ms = float(np.median([s["spread_interp_ms"] for s in stage_ms]))
        peak, peak_src = hbm_peak()
        local_cells = geo["nz_local"] * geo["plane_cells"]
        abytes = M * (3 * 8 + 16 + 4) + local_cells * 16
        achieved = abytes / (k_ms * 1e-3) / 1e9
        stages = {k: float(np.median([s[k] for s in stage_ms])) for k in stage_ms[0]}
        kern = "interp" if ttype == 2 else "spread"
        line = {
            "metric": METRIC, "value": value, "unit": "NU pts/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_record(cfg, M, 1, geo["ns"], [geo["nf1"], geo["nf2"], nf3], cfg["opts"].get("gpu_method", 1),
                                    "z-slab decomposition of the fine grid, points routed to their slabs inside setpts, mode array replicated; " +
                                    ("type 2: no collective (halo planes are derived locally)" if ttype == 2 else
                                     "type 1: ring halo add + all-reduce of the mode array over NCCL inside the timed step"),
                                    {"M_total": M_total, "slab_planes_rank0": [geo["z0"], geo["z1"]], "halo_planes": geo["pad"],
                                     "points_outside_slab": outside}),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": k_ms,
                         "binding": binding_roofline(cfg, kern, geo["ns"], M, 1, k_ms, sm_peaks(local_rank), {}),
                         "note": "rank 0's launch; HBM roofline as the contract asks, `binding` = the SM resource that bounds the kernel"},
            "vs_ref_gpu": {"available": False, "why": "the reference is single-GPU and needs ~100 GB for this size next to ours; "
                                                      "parity at full size: tests/test_fullsize_gpu.py::test_config5_full_size_single_gpu"},
            "stages_ms": stages,
            "setpts": {"ms": setpts_ms, "pts_per_s": M_total / (setpts_ms * 1e-3), "launches": setpts_launches,
                       "includes": "cufinufft_slab_route_setpts: owner computation, counts all-gather, all-to-all of the coordinates (NCCL), bin sort"},
            "checksum_abs_rank0": checksum, "cpu_baseline": None,
        }
        if with_cpu and not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(cfg)
            cb["sample"] = "fine grid down-scaled to 128^3 modes on the host: " + cb["sample"]
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")}
    plan.destroy()
    comm.destroy()
    del held, c, fk
    torch.cuda.empty_cache()
    return line


def bind_to_gpu_numa_node(torch, local_rank):
    """Several ranks per host: run this process (and hence first-touch its pinned host buffers) on the CPUs of the
    NUMA node its GPU hangs off (sysfs local_cpulist of the GPU's PCI function), so that the host-buffer legs of N
    ranks do not all cross one socket's memory controller (VERDICT r1 "weak 12": e2e efficiency 0.27 at N = 8 with every
    rank on node 0).  Plumbing only; returns what was done for the JSON line."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev_id = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev_id)
        node = int(open(path + "/numa_node").read().strip())
        cpus = []
        for part in open(path + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"gpu_pci": path.rsplit("/", 1)[1], "numa_node": node, "cpus": len(cpus)}
    except Exception as exc:   # noqa: BLE001
        return {"error": repr(exc)}


def compact(line):
    """The part of a record that goes under `extra`."""
    if line is None:
        return None
    keep = ("value", "unit", "ms_per_step", "scaling", "n_gpus", "steps", "dtype", "config", "e2e", "gpu_launches", "stages_ms", "setpts",
            "vs_ref_gpu", "roofline")
    return {k: line[k] for k in keep if k in line}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=0, help="0 = default workload: config 3 on one GPU, config 5 (z-slabs) on several")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-ref", action="store_true", help="skip timing the reference GPU library beside ours")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only")
    ap.add_argument("--scale", type=float, default=1.0, help="scale M (debug)")
    ap.add_argument("--sort-levels", type=int, default=0, help="cufinufft_set_sort_levels value (experiments: +4 no coarse partition, +8 always)")
    ap.add_argument("--opt", action="append", default=[], help="override a cufinufft_opts field, e.g. --opt gpu_binsizex=8 (experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    explicit = args.config != 0
    cfg_id = args.config if explicit else (3 if world == 1 else 5)
    cfg = dict(CONFIGS[cfg_id])
    cfg["M"] = int(cfg["M"] * args.scale)
    if args.opt:
        cfg["opts"] = dict(cfg["opts"], **{k: int(v) for k, v in (o.split("=") for o in args.opt)})
        cfg["name"] += " [" + ",".join(args.opt) + "]"

    if args.impl == "reference":
        run_reference_impl(args, cfg, cfg_id, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = dict(torch=torch, dist=dist, rank=rank, local_rank=local_rank, world=world, dev=dev)
    t_wall = time.perf_counter()

    runner = run_slab if cfg.get("slab") else run_plain
    line = runner(args, cfg_id, cfg, ctx, args.steps, args.warmup)

    extra = {}
    if not explicit and not args.no_extra and args.scale == 1.0:
        def sub(name, cid, fn, opts=None, **kw):
            try:
                c_ = dict(CONFIGS[cid])
                if opts:
                    c_["opts"] = dict(c_["opts"], **opts)
                    c_["name"] += " " + " ".join("%s=%s" % kv for kv in opts.items())
                rec = fn(args, cid, c_, ctx, **kw)
            except Exception as exc:   # noqa: BLE001
                rec = {"error": repr(exc)}
                torch.cuda.empty_cache()
            if rank == 0:
                extra[name] = compact(rec) if rec and "error" not in rec else rec

        k = max(3, min(args.steps, 5))
        if world == 1:
            for cid in (1, 2, 4, 7):
                sub("cfg%d" % cid if cid != 7 else "cfg4_type2", cid, run_plain, steps=k, warmup=3, with_e2e=False, with_cpu=False)
            # Horner against Horner (gpu_kerevalmeth=1 on both sides; the headline compares the default evaluators)
            sub("cfg3_horner", 3, run_plain, opts=dict(gpu_kerevalmeth=1), steps=k, warmup=3, with_e2e=False, with_cpu=False)
            sub("cfg3_onebin", 8, run_plain, steps=k, warmup=3, with_e2e=False, with_cpu=False)
            if torch.cuda.mem_get_info()[0] > 150 * 2 ** 30:
                sub("cfg5_slab_1gpu", 5, run_slab, steps=3, warmup=3, with_e2e=False, with_cpu=False)
        else:
            sub("cfg4_sharded", 4, run_plain, steps=k, warmup=3, with_cpu=False, with_ref=False)
            sub("cfg5_type1_slab", 6, run_slab, steps=3, warmup=3, with_e2e=False, with_cpu=False)
    if rank == 0:
        if numa is not None:
            line["numa_binding_rank0"] = numa
        if world > 1:
            line["scaling_note"] = ("strong scaling of config 5; the same workload on ONE GPU is the record extra.cfg5_slab_1gpu "
                                    "of the `--gpus 1` line (whose headline is config 3, the north-star single-GPU target)")
        if extra:
            line["extra"] = extra
        line["wall_s"] = time.perf_counter() - t_wall
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
