#!/bin/bash
OUT=gpurun_out/${1:-r2f}
mkdir -p $OUT
timeout 600 python bench.py --config 9 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ref > $OUT/bench_cfg9.json 2> $OUT/bench_cfg9.err; tail -c 900 $OUT/bench_cfg9.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:interp_tile_kernel -s 1 -c 1 -o $OUT/prof_interp_tile_cfg9 \
  python bench.py --config 9 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ref > $OUT/ncu_cfg9.log 2>&1; echo "rc=$?"
