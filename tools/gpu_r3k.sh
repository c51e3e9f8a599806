#!/bin/bash
OUT=gpurun_out/${1:-r3k}
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== racecheck plane engine"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 40 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "plane_owner and (uniform or onebin or wide)" > $OUT/racecheck_plane.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_plane.log | head
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "plane_owner" > $OUT/memcheck_plane.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/memcheck_plane.log | head
