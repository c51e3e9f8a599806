#!/bin/bash
OUT=gpurun_out/${1:-r3r}
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -i numa > $OUT/numa.txt
n=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --steps 5 --warmup 3 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "bench n=$n rc=$?"; tail -c 300 $OUT/bench_n$n.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n8.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f setpts %s e2e %s numa %s" % (d["value"], d["ms_per_step"], (d.get("setpts") or {}).get("ms"), (d.get("e2e") or {}).get("ms_per_step"), d.get("numa_binding_rank0")))
print("   stages", d.get("stages_ms"))
for k,v in (d.get("extra") or {}).items():
    print("   extra", k, (v or {}).get("value"), (v or {}).get("ms_per_step"), (v or {}).get("stages_ms"), ((v or {}).get("e2e") or {}).get("ms_per_step"), (v or {}).get("error"))
PY
cat $OUT/numa.txt
