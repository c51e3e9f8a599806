#!/bin/bash
OUT=gpurun_out/${1:-r2l}
mkdir -p $OUT
for k in coarse_count coarse_scatter; do
ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -o $OUT/prof_$k -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref --sort-levels 8 > $OUT/ncu_$k.log 2>&1
echo "$k rc=$?"
done
ls -la $OUT
