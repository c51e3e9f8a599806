#!/bin/bash
OUT=gpurun_out/${1:-r3i}
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"spread_plane" -s 1 -c 1 -o $OUT/prof_spread_plane_cfg10 -f python bench.py --config 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/ncu.log 2>&1
echo "rc=$?"; ls -la $OUT
