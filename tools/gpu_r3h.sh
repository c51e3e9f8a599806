#!/bin/bash
OUT=gpurun_out/${1:-r3h}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_gpu.log
for pl in 1 0; do
  CFB_PLANE=$pl timeout 600 python bench.py --config 10 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_cfg10_plane$pl.json 2> $OUT/bench_cfg10_plane$pl.err
  python -c "
import json;d=json.loads(open('$OUT/bench_cfg10_plane$pl.json').read().strip().splitlines()[-1]);v=d.get('vs_ref_gpu') or {};print('plane $pl exec', d['ms_per_step'], 'spread', d['stages_ms']['spread_interp_ms'], 'setpts', d['setpts']['ms'], 'ref', v.get('ref_exec_ms'), 'rel', v.get('rel_l2_ours_vs_ref'), v.get('why'))" || tail -3 $OUT/bench_cfg10_plane$pl.err
done
