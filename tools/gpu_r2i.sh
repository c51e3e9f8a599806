#!/bin/bash
OUT=gpurun_out/${1:-r2i}
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt
echo "== pytest full"; timeout 1500 python -m pytest tests -m gpu -q -x -rs --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench default N=1"; timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "rc=$?"; tail -c 300 $OUT/bench_default.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  echo "== bench N=$NG"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 5 --warmup 3 > $OUT/bench_n$NG.json 2> $OUT/bench_n$NG.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_n$NG.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms/step %.3f setpts %s e2e %s" % (d["value"], d["ms_per_step"], (d.get("setpts") or {}).get("ms"), (d.get("e2e") or {}).get("ms_per_step")))
        print("   roofline", d.get("roofline"))
        print("   vs_ref_gpu", d.get("vs_ref_gpu"))
        for k,v in (d.get("extra") or {}).items():
            print("   extra", k, (v or {}).get("value"), (v or {}).get("ms_per_step"), (v or {}).get("setpts"), (v or {}).get("e2e"), (v or {}).get("error"))
    except Exception as e: print(f, "unreadable", e)
PY
