#!/bin/bash
OUT=gpurun_out/${1:-r3t}
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"interp_tile" -s 1 -c 1 -o $OUT/prof_interp_tile_cfg9 -f python bench.py --config 9 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/ncu.log 2>&1
echo "rc=$?"
