#!/bin/bash
OUT=gpurun_out/${1:-r2j}
mkdir -p $OUT
echo "== pytest parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_vs_reference_gpu.py tests/test_slab_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_gpu.log
for c in 3 1 2; do
for sl in 4 8; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --sort-levels $sl > $OUT/bench_cfg${c}_sl$sl.json 2> $OUT/bench_cfg${c}_sl$sl.err; echo "cfg $c sl $sl rc=$?"
done; done
true
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"coarse|key_count|place_points|scan_|local_sort|map_sub|ref_bins|isub" -c 60 --csv --log-file $OUT/launches_setpts_cfg3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/ncu_bench.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f setpts %s" % (d["ms_per_step"], (d.get("setpts") or {})), "ref setpts", (d.get("vs_ref_gpu") or {}).get("ref_setpts_ms"))
    except Exception as e: print(f, "unreadable", e)
PY
