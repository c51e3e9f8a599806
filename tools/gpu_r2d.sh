#!/bin/bash
# round 2, call D: sm2 engine tuning sweep on config 3 (sub-bin target, work-item cap) + tests
OUT=gpurun_out/${1:-r2d}
mkdir -p $OUT
echo "== pytest gpu (spread-related)"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --config ${CFG:-3} --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("$name: spread %.3f ms  total %.3f  setpts %.3f" % (d["stages_ms"]["spread_interp_ms"], d["ms_per_step"], d["setpts"]["ms"]))
except Exception as e:
    print("$name failed", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
}
run base X=1
run t20 CFB_SM_TARGET_KB=20
run t24 CFB_SM_TARGET_KB=24
run t28 CFB_SM_TARGET_KB=28
run t40 CFB_SM_TARGET_KB=40
run t10 CFB_SM_TARGET_KB=10
run item8k CFB_SM_MAXITEM=8192
run item16k_t24 CFB_SM_MAXITEM=16384 CFB_SM_TARGET_KB=24
CFG=1 run cfg1 X=1
CFG=4 run cfg4 X=1
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spread_sm2_kernel -s 1 -c 1 -o $OUT/prof_spread_sm2_cfg3 \
  python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg3.log 2>&1; echo "rc=$?"
