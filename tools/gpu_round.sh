#!/bin/bash
# tools/gpu_round.sh -- one gpurun call: smoke, GPU parity tests, bench lines, reference
# comparison, ncu launch list + one full capture of the top kernel.  Everything lands in
# gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by hand.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI_PID=$!

echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log

for c in 2 1 3 4; do
  echo "== bench cfg $c"
  extra="--no-cpu-baseline"; [ $c = 2 ] && extra=""
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 $extra > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err
  echo "rc=$?"; tail -c 1500 $OUT/bench_cfg$c.json
done
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json

echo "== compare vs reference library"
timeout 900 python tools/compare_reference.py --configs 1,2,3,4 --reps 3 > $OUT/compare.jsonl 2> $OUT/compare.err; echo "rc=$?"; cat $OUT/compare.jsonl

kill $SMI_PID

echo "== ncu launch lists"
for c in 2 1 3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg$c.csv \
    python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_cfg$c.log 2>&1
  echo "cfg $c rc=$?"
done
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 1 -c 1 -o $OUT/prof_interp_cfg2 \
  python bench.py --config 2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spread_sm_kernel -s 1 -c 1 -o $OUT/prof_spread_cfg1 \
  python bench.py --config 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg1.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spread_sm_kernel -s 1 -c 1 -o $OUT/prof_spread_cfg3 \
  python bench.py --config 3 --scale 0.25 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_cfg3.log 2>&1; echo "rc=$?"
ls -la $OUT
