#!/bin/bash
# tools/gpu_multi.sh <tag> "<N list>" [test] -- multi-GPU round on one box: optional NCCL slab test, then for every N
# bench lines of config 5 (z-slab, strong), config 4 (batch sharded by transform, strong) and the default
# config 2 (independent transforms, weak), launched as the driver does (torchrun, one rank per GPU).
TAG=${1:-multi}
NS=${2:-"2"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "$3" = "test" ]; then
  echo "== NCCL slab test"; timeout 600 python -m pytest tests/test_slab_gpu.py -q -k nccl > $OUT/pytest_nccl.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_nccl.log
fi
P=29500
for N in $NS; do
  for C in 5 4 2; do
    P=$((P+1))
    extra="--no-cpu-baseline"
    echo "== bench cfg $C N=$N"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --config $C --steps 5 --warmup 3 $extra > $OUT/bench_cfg${C}_n$N.json 2> $OUT/bench_cfg${C}_n$N.err
    echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/bench_cfg${C}_n$N.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, d["stages_ms"], (d["e2e"] or {}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
    tail -2 $OUT/bench_cfg${C}_n$N.err
  done
done
