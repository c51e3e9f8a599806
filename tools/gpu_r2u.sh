#!/bin/bash
OUT=gpurun_out/${1:-r2u}
mkdir -p $OUT
for v in "8 8" "4 8" "16 8" "8 4" "8 2" "4 4"; do
  set -- $v
  touch cufinufft_b200/csrc/setpts.cu
  make -C cufinufft_b200/csrc EXTRA="-DCFB_CP_PPT=$1 -DCFB_CP_BPS=$2" -j8 > $OUT/make_$1_$2.log 2>&1 || { echo "make failed $v"; tail -5 $OUT/make_$1_$2.log; continue; }
  timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref --sort-levels 8 > $OUT/bench_cfg3_ppt$1_bps$2.json 2> $OUT/bench_cfg3_ppt$1_bps$2.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg3_ppt$1_bps$2.json').read().strip().splitlines()[-1]);print('ppt $1 bps $2 setpts', d['setpts']['ms'])"
done
