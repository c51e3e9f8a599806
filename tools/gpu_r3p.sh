#!/bin/bash
OUT=gpurun_out/${1:-r3p}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
for c in 9 2; do
timeout 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
python -c "
import json;d=json.loads(open('$OUT/bench_cfg$c.json').read().strip().splitlines()[-1]);print('cfg $c exec', d['ms_per_step'], 'interp', d['stages_ms']['spread_interp_ms'])" || tail -3 $OUT/bench_cfg$c.err
done
