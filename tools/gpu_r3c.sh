#!/bin/bash
OUT=gpurun_out/${1:-r3c}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "rc=$?"; tail -c 300 $OUT/bench_default.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f setpts %s" % (d["value"], d["ms_per_step"], d.get("setpts")))
print("e2e", d.get("e2e"))
print("roofline", {k:v for k,v in d["roofline"].items() if k!='binding'})
print("binding", [(v['resource'], round(v['frac'],3)) for v in d["roofline"]["binding"]["views"]])
print("vs_ref_gpu", d.get("vs_ref_gpu"))
print("cpu_baseline", d.get("cpu_baseline"))
for k,v in (d.get("extra") or {}).items():
    print("   extra", k, "value %s ms %s setpts %s e2e %s vsref %s" % ((v or {}).get("value"), (v or {}).get("ms_per_step"), ((v or {}).get("setpts") or {}).get("ms"), ((v or {}).get("e2e") or {}).get("ms_per_step"), {kk:vv for kk,vv in ((v or {}).get("vs_ref_gpu") or {}).items() if 'speed' in kk or 'ref_' in kk}), (v or {}).get("error"))
PY
