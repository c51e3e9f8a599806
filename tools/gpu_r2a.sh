#!/bin/bash
# round 2, call A: the whole GPU test suite (new parity tests included) + baseline bench lines
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1; nproc > $OUT/nproc.txt; free -g > $OUT/mem.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -rs --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_gpu.log
for c in 3 1; do
  echo "== bench cfg $c"
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
  echo "rc=$?"; tail -c 1200 $OUT/bench_cfg$c.json
done
