#!/bin/bash
OUT=gpurun_out/${1:-r2z}
mkdir -p $OUT
nvidia-smi -L > $OUT/smi.txt; NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest mgpu + slab"; timeout 900 python -m pytest tests/test_mgpu_gpu.py tests/test_slab_gpu.py -m gpu -q -rs > $OUT/pytest_mgpu.log 2>&1; echo "rc=$?"; tail -6 $OUT/pytest_mgpu.log
for n in 8 4 2; do
  [ "$NG" -ge $n ] || continue
  X=""; [ $n -ne 8 ] && X="--no-extra"
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 $X > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "bench n=$n rc=$?"; tail -c 300 $OUT/bench_n$n.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e ms/step %.3f setpts %s e2e %s" % (d["value"], d["ms_per_step"], (d.get("setpts") or {}).get("ms"), (d.get("e2e") or {}).get("ms_per_step")))
        print("   stages", d.get("stages_ms"))
        for k,v in (d.get("extra") or {}).items():
            print("   extra", k, (v or {}).get("value"), (v or {}).get("ms_per_step"), (v or {}).get("stages_ms"), (v or {}).get("error"))
    except Exception as e: print(f, "unreadable", e)
PY
