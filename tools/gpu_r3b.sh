#!/bin/bash
OUT=gpurun_out/${1:-r3b}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
for zm in 4 3 2 1; do
  CFB_INTERP_ZM=$zm timeout 300 python bench.py --config 9 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/bench_cfg9_zm$zm.json 2> $OUT/bench_cfg9_zm$zm.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg9_zm$zm.json').read().strip().splitlines()[-1]);print('zm $zm exec', d['ms_per_step'], 'interp', d['stages_ms']['spread_interp_ms'])" || tail -3 $OUT/bench_cfg9_zm$zm.err
done
