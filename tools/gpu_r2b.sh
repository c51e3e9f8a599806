#!/bin/bash
# round 2, call B: new SM spread engine -- tests, A/B timing of configs 3, 1, 4 (gen 1 vs gen 2), launch list + one full ncu capture
OUT=gpurun_out/${1:-r2b}
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
for c in 3 1 4; do
  for g in 2 1; do
    echo "== bench cfg $c gen $g"
    CFB_SM_GEN=$g timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_cfg${c}_gen$g.json 2> $OUT/bench_cfg${c}_gen$g.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_cfg${c}_gen$g.json").read().strip().splitlines()[-1])
    print("cfg $c gen $g: ms/step %.3f  stages %s  setpts %.3f ms" % (d["ms_per_step"], {k: round(v,3) for k,v in d["stages_ms"].items()}, d["setpts"]["ms"]))
except Exception as e:
    print("failed", e); print(open("$OUT/bench_cfg${c}_gen$g.err").read()[-2000:])
PY
  done
done
echo "== ncu full capture cfg3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spread_sm2_kernel -s 1 -c 1 -o $OUT/prof_spread_sm2_cfg3 \
  python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg3.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spread_sm2_kernel -s 1 -c 1 -o $OUT/prof_spread_sm2_cfg1 \
  python bench.py --config 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg1.log 2>&1; echo "rc=$?"
ls -la $OUT
