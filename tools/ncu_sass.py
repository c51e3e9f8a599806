#!/usr/bin/env python
"""tools/ncu_sass.py <rep> [out.txt] -- per-SASS-instruction executed counts and stall samples of one
ncu --set full --import-source on capture (the source page), as a compact listing."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iE, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
tot_e = sum(int(r[iE]) for r in rows[2:] if len(r) > iE and r[iE].isdigit())
tot_s = sum(int(r[iN]) for r in rows[2:] if len(r) > iN and r[iN].isdigit())
f = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
f.write("# total inst executed %d, samples %d\n#  idx   exec%%  samp%%  thr  SASS\n" % (tot_e, tot_s))
for i, r in enumerate(rows[2:]):
    if len(r) <= iE or not r[iE].isdigit():
        continue
    f.write("%5d %6.2f %6.2f %4s  %s\n" % (i, 100.0 * int(r[iE]) / max(tot_e, 1), 100.0 * int(r[iN]) / max(tot_s, 1), r[iT], r[iS].strip()))
