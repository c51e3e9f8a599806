#!/bin/bash
# tools/gpu_sort2.sh <tag> -- two-level point order: tests, then setpts/exec of configs 3 and 5 with both modes
TAG=${1:-sort2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for c in 3 1 2; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
  python - <<PY
import json
d = json.loads(open("$OUT/bench_cfg$c.json").read().strip().splitlines()[-1])
print("cfg$c exec %.3f ms  spread/interp %.3f  setpts %.3f ms (%d launches)" % (d["ms_per_step"], d["stages_ms"]["spread_interp_ms"], d["setpts"]["ms"], d["setpts"]["launches"]))
PY
done
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5 exec %.3f ms  interp %.3f  setpts %.3f ms" % (d["ms_per_step"], d["stages_ms"]["spread_interp_ms"], d["setpts"]["ms"]))
PY
