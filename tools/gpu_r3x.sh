#!/bin/bash
OUT=gpurun_out/${1:-r3x}
mkdir -p $OUT
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"spread_sm2" -s 3 -c 1 -o $OUT/prof_spread_sm2_cfg3 -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/ncu_spread.log 2>&1
echo "spread rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"interp_tile" -s 1 -c 1 -o $OUT/prof_interp_tile_cfg9 -f python bench.py --config 9 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/ncu_interp.log 2>&1
echo "interp rc=$?"
