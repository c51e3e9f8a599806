#!/bin/bash
OUT=gpurun_out/${1:-r2g}
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for c in 9 2 7; do
timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-ref > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg$c.json").read().strip().splitlines()[-1])
    print("cfg $c: interp %.3f ms  total %.3f  setpts %.3f" % (d["stages_ms"]["spread_interp_ms"], d["ms_per_step"], d["setpts"]["ms"]))
except Exception as e:
    print("cfg $c failed", e); print(open("$OUT/bench_cfg$c.err").read()[-1500:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:interp_tile_kernel -s 1 -c 1 -o $OUT/prof_interp_tile_cfg9 \
  python bench.py --config 9 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ref > $OUT/ncu_cfg9.log 2>&1; echo "rc=$?"
