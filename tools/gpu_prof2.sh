#!/bin/bash
# tools/gpu_prof2.sh <tag> "<cfg:kernel-regex:scale> ..." -- ncu --set full + source of one launch per entry
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for e in $2; do
  IFS=: read c k s <<< "$e"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/prof_${k}_cfg$c \
    python bench.py --config $c --scale ${s:-1.0} --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_${k}_cfg$c.log 2>&1; echo "cfg $c $k rc=$?"
done
ls -la $OUT
