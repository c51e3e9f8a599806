#!/bin/bash
OUT=gpurun_out/${1:-r3q}
mkdir -p $OUT
for k in 1 2 3; do
CFB_PIPE3D=$k timeout 300 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-ref > $OUT/bench_cfg3_k$k.json 2> $OUT/bench_cfg3_k$k.err
python -c "
import json;d=json.loads(open('$OUT/bench_cfg3_k$k.json').read().strip().splitlines()[-1]);print('K=$k exec', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" || tail -3 $OUT/bench_cfg3_k$k.err
done
