#!/bin/bash
OUT=gpurun_out/${1:-r3e}
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
for c in 3 1 2; do
timeout 300 python bench.py --config $c --opt gpu_kerevalmeth=1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_cfg${c}_horner.json 2> $OUT/bench_cfg${c}_horner.err
python -c "
import json;d=json.loads(open('$OUT/bench_cfg${c}_horner.json').read().strip().splitlines()[-1]);v=d['vs_ref_gpu'];print('cfg $c horner exec', d['ms_per_step'], 'ref', v.get('ref_exec_ms'), 'speedup', v.get('speedup_exec'), 'rel', v.get('rel_l2_ours_vs_ref'))" || tail -3 $OUT/bench_cfg${c}_horner.err
done
