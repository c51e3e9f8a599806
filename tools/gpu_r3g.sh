#!/bin/bash
OUT=gpurun_out/${1:-r3g}
mkdir -p $OUT
run() { tag=$1; shift
  timeout 600 python bench.py --config 10 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref "$@" > $OUT/bench_cfg10_$tag.json 2> $OUT/bench_cfg10_$tag.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg10_$tag.json').read().strip().splitlines()[-1]);print('$tag exec', d['ms_per_step'], 'spread', d['stages_ms']['spread_interp_ms'], 'setpts', d['setpts']['ms'])" || tail -3 $OUT/bench_cfg10_$tag.err
}
run base
run x8y8 --opt gpu_binsizex=8 --opt gpu_binsizey=8
run x4y4 --opt gpu_binsizex=4 --opt gpu_binsizey=4
run x4y4z4 --opt gpu_binsizex=4 --opt gpu_binsizey=4 --opt gpu_binsizez=4
run x2y2 --opt gpu_binsizex=2 --opt gpu_binsizey=2
run gm --opt gpu_method=1
