#!/bin/bash
# tools/gpu_quick.sh -- short GPU iteration: parity tests, device-resident timings of the four
# single-GPU configs against the reference library, accuracy audit.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 900 python tools/compare_reference.py --configs ${2:-1,2,3,4} --reps 5 > $OUT/compare.jsonl 2> $OUT/compare.err; echo "compare rc=$?"; cat $OUT/compare.jsonl; tail -5 $OUT/compare.err
timeout 600 python tools/check_accuracy.py --configs ${2:-1,2,3,4} > $OUT/accuracy.jsonl 2> $OUT/accuracy.err; echo "accuracy rc=$?"; cat $OUT/accuracy.jsonl; tail -5 $OUT/accuracy.err
