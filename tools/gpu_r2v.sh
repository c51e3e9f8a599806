#!/bin/bash
OUT=gpurun_out/${1:-r2v}
mkdir -p $OUT
for c in 9 8 3 2 1; do for sl in 4 8; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref --sort-levels $sl > $OUT/bench_cfg${c}_sl$sl.json 2> $OUT/bench_cfg${c}_sl$sl.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg${c}_sl$sl.json').read().strip().splitlines()[-1]);print('cfg $c sl $sl setpts', d['setpts']['ms'], 'exec', d['ms_per_step'])"
done; done
