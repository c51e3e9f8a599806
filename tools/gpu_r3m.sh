#!/bin/bash
OUT=gpurun_out/${1:-r3m}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err
python -c "
import json;d=json.loads(open('$OUT/bench_cfg3.json').read().strip().splitlines()[-1]);print('cfg 3 exec', d['ms_per_step'], 'spread', d['stages_ms']['spread_interp_ms'])" || tail -3 $OUT/bench_cfg3.err
ncu --set full --clock-control none --import-source on -k regex:"spread_sm2" -s 3 -c 1 -o $OUT/prof_spread_sm2_cfg3 -f python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/ncu.log 2>&1
echo "ncu rc=$?"
