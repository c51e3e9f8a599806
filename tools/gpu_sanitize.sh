#!/bin/bash
# tools/gpu_sanitize.sh <tag> -- compute-sanitizer memcheck over a spread of small GPU tests, racecheck on two of them
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
echo "== memcheck"
SEL=""
for i in 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do SEL="$SEL tests/test_gpu_random_configs.py::test_random_plan_matches_oracle[$i]"; done
SEL="$SEL tests/test_slab_gpu.py::test_emulated_slabs_match_undivided_plan[24x20x32-float32-uniform-2] tests/test_slab_gpu.py::test_emulated_slabs_match_undivided_plan[20x18x32-float64-wide-3] tests/test_slab_gpu.py::test_points_outside_the_slab_are_counted"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest $SEL -q -x > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck.log | head -20
echo "== racecheck (every engine: the fixed parity cases, sub-bin and two-level cases)"
timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 40 \
  python -m pytest tests/test_gpu_parity.py -q -k "test_transform_vs_oracle or (two_level and (t1-128x96 or t2-64x48 or t1-32x32x32)) or (subbins and 20x18x16)" > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck.log | head; grep -E "Race reported|hazards\]" $OUT/racecheck.log | sed -E 's/\(cfb::SIArgs<T1>\).*operator \(\)[^+]*//' | cut -c1-200 | sort | uniq -c | sort -rn | head -30
