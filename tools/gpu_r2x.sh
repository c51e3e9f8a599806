#!/bin/bash
OUT=gpurun_out/${1:-r2x}
mkdir -p $OUT
for c in 2 3 1; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-extra --no-ref > $OUT/bench_cfg${c}.json 2> $OUT/bench_cfg${c}.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg${c}.json').read().strip().splitlines()[-1]);print('cfg $c setpts', d['setpts']['ms'], 'exec', d['ms_per_step'], 'e2e', d['e2e'])"
  tail -3 $OUT/bench_cfg${c}.err
done
