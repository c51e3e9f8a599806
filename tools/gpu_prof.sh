#!/bin/bash
# tools/gpu_prof.sh <tag> [cfg3scale] -- ncu --set full of the top kernel of configs 1,2,3 (one launch each)
TAG=${1:-prof}
S3=${2:-1.0}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 1 -c 1 -o $OUT/prof_interp_cfg2 \
  python bench.py --config 2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spread_sm_kernel -s 1 -c 1 -o $OUT/prof_spread_cfg1 \
  python bench.py --config 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg1.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spread_sm_kernel -s 1 -c 1 -o $OUT/prof_spread_cfg3 \
  python bench.py --config 3 --scale $S3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg3.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg3.csv \
    python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_cfg3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg2.csv \
    python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_cfg2.log 2>&1
ls -la $OUT
