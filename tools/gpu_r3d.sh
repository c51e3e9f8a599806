#!/bin/bash
OUT=gpurun_out/${1:-r3d}
mkdir -p $OUT
for c in 8; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "rc=$?"
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg$c.json').read().strip().splitlines()[-1]);print('cfg $c exec', d['ms_per_step'], 'setpts', d['setpts']['ms'], d['vs_ref_gpu'])" || tail -3 $OUT/bench_cfg$c.err
done
timeout 300 python bench.py --config 3 --opt gpu_kerevalmeth=1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_cfg3_horner.json 2> $OUT/bench_cfg3_horner.err
python -c "
import json;d=json.loads(open('$OUT/bench_cfg3_horner.json').read().strip().splitlines()[-1]);print('cfg 3 horner exec', d['ms_per_step'], 'setpts', d['setpts']['ms'], d['vs_ref_gpu'])" || tail -3 $OUT/bench_cfg3_horner.err
