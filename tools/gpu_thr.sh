#!/bin/bash
# tools/gpu_thr.sh <tag> -- sweep of the short-run threshold (CFB_DIRECT_THR) on configs 1, 3, 4
TAG=${1:-thr}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for thr in "" 2/1 3/1 4/1 6/1; do
  for c in 1 4; do
    CFB_DIRECT_THR=$thr timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/b.json 2> $OUT/b.err
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
    print("thr=[$thr] cfg$c exec %.3f ms spread %.3f" % (d["ms_per_step"], d["stages_ms"]["spread_interp_ms"]))
except Exception as e:
    print("thr=[$thr] cfg$c failed", e)
PY
  done
done
