#!/bin/bash
# tools/gpu_quick2.sh <tag> -- GPU tests + type-1 slab bench (config 6, 1/8 size) + default bench line
TAG=${1:-q2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== bench cfg6 x0.125"; timeout 600 python bench.py --config 6 --scale 0.125 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_cfg6_eighth.json 2> $OUT/bench_cfg6_eighth.err; echo "rc=$?"; tail -c 1800 $OUT/bench_cfg6_eighth.json; tail -3 $OUT/bench_cfg6_eighth.err
echo "== bench default"; timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "rc=$?"; tail -c 2800 $OUT/bench_default.json; tail -3 $OUT/bench_default.err
