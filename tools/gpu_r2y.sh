#!/bin/bash
OUT=gpurun_out/${1:-r2y}
mkdir -p $OUT
for v in 130 140 170 215; do
  touch cufinufft_b200/csrc/spread.cu
  make -C cufinufft_b200/csrc EXTRA="-DCFB_TILE_PAD_PCT=$v" -j8 > $OUT/make_$v.log 2>&1 || { echo "make failed $v"; tail -5 $OUT/make_$v.log; continue; }
  for c in 3 1; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/bench_cfg${c}_pad$v.json 2> $OUT/bench_cfg${c}_pad$v.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg${c}_pad$v.json').read().strip().splitlines()[-1]);print('pad $v cfg $c exec', d['ms_per_step'], 'spread', d['stages_ms']['spread_interp_ms'], 'binding', [(x['resource'], round(x['frac'],3)) for x in d['roofline']['binding']['views']])"
  done
done
