#!/bin/bash
OUT=gpurun_out/${1:-r3a}
mkdir -p $OUT
run() { tag=$1; shift
  timeout 300 python bench.py --config 9 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extra --no-ref "$@" > $OUT/bench_cfg9_$tag.json 2> $OUT/bench_cfg9_$tag.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg9_$tag.json').read().strip().splitlines()[-1]);print('$tag exec', d['ms_per_step'], 'interp', d['stages_ms']['spread_interp_ms'], 'setpts', d['setpts']['ms'])" || tail -3 $OUT/bench_cfg9_$tag.err
}
run base
run y8 --opt gpu_binsizey=8
run x8y8 --opt gpu_binsizex=8 --opt gpu_binsizey=8
run y8z4 --opt gpu_binsizey=8 --opt gpu_binsizez=4
run x8y8z4 --opt gpu_binsizex=8 --opt gpu_binsizey=8 --opt gpu_binsizez=4
run z4 --opt gpu_binsizez=4
