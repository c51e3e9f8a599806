#!/bin/bash
# tools/gpu_check.sh [tag] [configs] -- correctness + timing round without ncu: smoke, GPU parity tests,
# bench lines of the single-GPU configs, comparison with the reference library, accuracy audit.
TAG=${1:-chk}
CFGS=${2:-"2 1 3 4"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
for c in $CFGS; do
  echo "== bench cfg $c"
  extra="--no-cpu-baseline"; [ $c = 2 ] && extra=""
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 $extra > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_cfg$c.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "stages_ms")}, d["e2e"].get("value"), d["roofline"]["frac"], d["setpts"]["ms"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
done
echo "== compare vs reference library"
timeout 900 python tools/compare_reference.py --configs $(echo $CFGS | tr ' ' ',') --reps 3 > $OUT/compare.jsonl 2> $OUT/compare.err; echo "rc=$?"
python - <<PY
import json
for l in open("$OUT/compare.jsonl"):
    d = json.loads(l)
    print(d["config"][:5], "ours set/exec %.3f/%.3f ref %.3f/%.3f  x%.2f  rel_l2 %s" % (d["ours_setpts_ms"], d["ours_exec_ms"], d["ref_setpts_ms"], d["ref_exec_ms"], d["speedup_exec"], d["rel_l2_vs_ref"]))
PY
tail -3 $OUT/compare.err
timeout 600 python tools/check_accuracy.py --configs $(echo $CFGS | tr ' ' ',') > $OUT/accuracy.jsonl 2> $OUT/accuracy.err; echo "accuracy rc=$?"; cat $OUT/accuracy.jsonl; tail -3 $OUT/accuracy.err
