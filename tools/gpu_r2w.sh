#!/bin/bash
OUT=gpurun_out/${1:-r2w}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
for c in 2 1 3; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_cfg${c}.json 2> $OUT/bench_cfg${c}.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg${c}.json').read().strip().splitlines()[-1]);print('cfg $c setpts', d['setpts']['ms'], 'exec', d['ms_per_step'], 'e2e', d['e2e'], 'vsref', {k:v for k,v in d['vs_ref_gpu'].items() if 'ms' in k or 'speed' in k})"
done
