#!/usr/bin/env python
"""tools/ncu_summary.py -- turn gpurun_out/<tag>/*.ncu-rep and launches_*.csv into the small text
summaries kept under profiles/ (the .ncu-rep files themselves are scratch).
  python tools/ncu_summary.py gpurun_out/r01a profiles/r01a"""
import collections
import csv
import glob
import os
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum")
STALL = "smsp__average_warps_issue_stalled_"


def rep_summary(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return "no data in %s\n" % path
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append("kernel: %s   (ncu --set full --clock-control none; one launch, cold caches)" % d.get("Kernel Name", "?"))
        for k in KEEP:
            if k in d:
                lines.append("  %-92s %14s %s" % (k, d[k], u[k]))
        stalls = sorted(((float(d[k].replace(",", "")), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k in d
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
        lines.append("  warp stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))
    return "\n".join(lines) + "\n"


def launch_summary(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        val = float(r[-1].replace(",", ""))
        val = val / 1e6 if r[-2] == "ns" else (val / 1e3 if r[-2] in ("us", "usecond") else val)
        a = agg.setdefault(r[4], [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for k, v in agg.items() if "cfb::" in k or "fft" in k)
    lines = ["%s  (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold-cache: compare SHARES)" % os.path.basename(path),
             "  share = of the library's own launches (cfb::* + cuFFT); torch kernels are the synthetic-input generation"]
    for k, v in agg.items():
        share = "%5.1f%%" % (100 * v[1] / tot) if ("cfb::" in k or "fft" in k) and tot else "      "
        lines.append("  n=%3d total %9.3f ms avg %8.3f ms %s  %s" % (v[0], v[1], v[1] / v[0], share, k[:110]))
    return "\n".join(lines) + "\n"


def main():
    src, dst = sys.argv[1], sys.argv[2]
    with open(dst + "_ncu_summary.txt", "w") as f:
        for p in sorted(glob.glob(os.path.join(src, "launches_*.csv"))):
            f.write(launch_summary(p) + "\n")
        for p in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
            f.write("== %s\n" % os.path.basename(p))
            f.write(rep_summary(p) + "\n")
    print(open(dst + "_ncu_summary.txt").read())


if __name__ == "__main__":
    main()
