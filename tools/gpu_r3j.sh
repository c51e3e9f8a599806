#!/bin/bash
OUT=gpurun_out/${1:-r3j}
mkdir -p $OUT
for pb in 256 64 128; do
  touch cufinufft_b200/csrc/spread.cu cufinufft_b200/csrc/si_f64_d3.cu
  make -C cufinufft_b200/csrc EXTRA="-DCFB_PLANE_PB=$pb -DCFB_DEV_NS" -j8 > $OUT/make_$pb.log 2>&1 || { echo "make failed $pb"; tail -5 $OUT/make_$pb.log; continue; }
  timeout 600 python bench.py --config 10 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-ref > $OUT/bench_cfg10_pb$pb.json 2> $OUT/bench_cfg10_pb$pb.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg10_pb$pb.json').read().strip().splitlines()[-1]);print('pb $pb exec', d['ms_per_step'], 'spread', d['stages_ms']['spread_interp_ms'])" || tail -3 $OUT/bench_cfg10_pb$pb.err
done
