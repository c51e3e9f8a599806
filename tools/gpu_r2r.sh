#!/bin/bash
OUT=gpurun_out/${1:-r2r}
mkdir -p $OUT
echo "== pytest parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_vs_reference_gpu.py tests/test_slab_gpu.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
for mb in 2 4 8; do
  CFB_SORT_BUCKET_MB=$mb timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref --sort-levels 8 > $OUT/bench_cfg3_mb$mb.json 2> $OUT/bench_cfg3_mb$mb.err; echo "mb $mb rc=$?"
done
for c in 1 2; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref --sort-levels 8 > $OUT/bench_cfg${c}_sl8.json 2> $OUT/bench_cfg${c}_sl8.err
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"coarse|key_count|place_points" -c 24 --csv --log-file $OUT/launches_setpts_cfg3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --no-ref --sort-levels 8 > $OUT/ncu_bench.log 2>&1
python - <<PY
import json,glob,csv,collections
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f setpts %s" % (d["ms_per_step"], (d.get("setpts") or {}).get("ms")))
    except Exception as e: print(f, "unreadable", e)
rows=list(csv.reader(open('$OUT/launches_setpts_cfg3.csv')))
hdr=None
agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        n=d['Kernel Name'][:50]+"|"+d['Metric Name']
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        agg.setdefault(n,[]).append((v,d['Metric Unit']))
for n,v in agg.items():
    print("%-90s n=%d  last=%.3f %s"%(n,len(v),v[-1][0],v[-1][1]))
PY
