#!/bin/bash
OUT=gpurun_out/${1:-r2h}
mkdir -p $OUT
echo "== pytest mgpu"; timeout 900 python -m pytest tests/test_mgpu_gpu.py tests/test_slab_gpu.py -m gpu -q -x -rs > $OUT/pytest_mgpu.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_mgpu.log
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
if [ "$NG" -ge 2 ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 5 --warmup 3 > $OUT/bench_n$NG.json 2> $OUT/bench_n$NG.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_n$NG.err
else
  timeout 900 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_n1.json 2> $OUT/bench_cfg5_n1.err; echo "bench rc=$?"; tail -c 400 $OUT/bench_cfg5_n1.err
fi
python - <<PY
import json,glob
for f in glob.glob("$OUT/bench_*.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms/step %.2f setpts %s e2e %s" % (d["value"], d["ms_per_step"], d["setpts"], (d.get("e2e") or {}).get("ms_per_step")))
        for k,v in (d.get("extra") or {}).items():
            print("   extra", k, (v or {}).get("value"), (v or {}).get("ms_per_step"), (v or {}).get("error"))
    except Exception as e: print(f, "unreadable", e)
PY
