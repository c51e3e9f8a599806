#!/bin/bash
# tools/gpu_slab.sh [tag] -- z-slab round on ONE GPU: slab tests (all ranks emulated on the device), the whole
# GPU suite, bench of config 5 at 1/8 and full size through the slab path, default bench line.
TAG=${1:-slab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
echo "== slab tests"; timeout 900 python -m pytest tests/test_slab_gpu.py -x -q > $OUT/pytest_slab.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_slab.log
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== bench cfg5 x0.125"; timeout 600 python bench.py --config 5 --scale 0.125 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg5_eighth.json 2> $OUT/bench_cfg5_eighth.err; echo "rc=$?"; tail -c 1500 $OUT/bench_cfg5_eighth.json; tail -3 $OUT/bench_cfg5_eighth.err
echo "== bench cfg5 full"; timeout 900 python bench.py --config 5 --steps 3 --warmup 3 > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "rc=$?"; tail -c 2500 $OUT/bench_cfg5.json; tail -3 $OUT/bench_cfg5.err
echo "== bench default"; timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "rc=$?"; tail -c 2500 $OUT/bench_default.json; tail -3 $OUT/bench_default.err
