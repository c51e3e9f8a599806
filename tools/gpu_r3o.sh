#!/bin/bash
OUT=gpurun_out/${1:-r3o}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_vs_reference_gpu.py tests/test_gpu_random_configs.py tests/test_fullsize_gpu.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
for c in 3 1 4; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
python -c "
import json;d=json.loads(open('$OUT/bench_cfg$c.json').read().strip().splitlines()[-1]);print('cfg $c exec', d['ms_per_step'], 'spread', d['stages_ms']['spread_interp_ms'])" || tail -3 $OUT/bench_cfg$c.err
done
