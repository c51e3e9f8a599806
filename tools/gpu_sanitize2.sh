#!/bin/bash
# tools/gpu_sanitize2.sh <tag> -- round 2: compute-sanitizer memcheck + racecheck over the kernels added or rewritten in round 2
TAG=${1:-san2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== memcheck"
SEL=""
for i in 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18 19; do SEL="$SEL tests/test_gpu_random_configs.py::test_random_plan_matches_oracle[$i]"; done
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest $SEL \
   tests/test_gpu_parity.py -q -x -k "random_plan or coarse_partitioned or two_level or transform_vs_oracle or subbins" > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck.log | head -20
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_slab_gpu.py tests/test_graph_capture_gpu.py tests/test_stage_parity_gpu.py -q -x -m gpu -k "not nccl" > $OUT/memcheck2.log 2>&1
echo "memcheck2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck2.log | head -20
echo "== racecheck: 2nd-gen SM spread, coarse partition, merged-tile interpolation, bank-class order"
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 40 \
  python -m pytest tests/test_gpu_parity.py tests/test_slab_gpu.py -q -m gpu -k "test_transform_vs_oracle or (coarse_partitioned and (t1-128x96 or t2-64x48 or t1-32x32x32 or t2-20x18x16)) or (subbins and 20x18x16) or (emulated_slabs and float64-wide-2)" > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck.log | head; grep -E "Race reported|hazards\]" $OUT/racecheck.log | sed -E 's/\(cfb::SIArgs<T1>\).*operator \(\)[^+]*//' | cut -c1-200 | sort | uniq -c | sort -rn | head -30
