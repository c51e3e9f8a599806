#!/bin/bash
# tools/gpu_final.sh <tag> -- GPU suite + the profiling evidence the bench line refers to:
#   launch list of the DEFAULT bench command (per-kernel shares), ncu --set full of the dominant kernels
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== launch list (default bench)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_cfg2.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_cfg3.csv \
    python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_cfg3.log 2>&1; echo "rc=$?"
for e in 2:interp_tile_kernel 3:spread_sm_kernel 1:spread_sm_kernel; do
  IFS=: read c k <<< "$e"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/prof_${k}_cfg$c \
    python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_${k}_cfg$c.log 2>&1; echo "cfg $c $k rc=$?"
done
ls -la $OUT
echo "== config 4, type-2 half (config 7): bench + reference comparison"
timeout 300 python bench.py --config 7 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg7.json 2> $OUT/bench_cfg7.err; echo "rc=$?"
timeout 300 python tools/compare_reference.py --configs 7 --reps 3 > $OUT/compare_cfg7.jsonl 2> $OUT/compare_cfg7.err; echo "rc=$?"; cat $OUT/compare_cfg7.jsonl
