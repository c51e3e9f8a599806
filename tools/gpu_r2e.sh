#!/bin/bash
OUT=gpurun_out/${1:-r2e}
mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
python -c "
import ctypes
from cufinufft_b200 import _cufinufft as ll
for w,n in enumerate(['lds128 B/s','ffma2 FMA/s','ffma FMA/s','dfma FMA/s','atoms /s','lds64 B/s']):
    o=ctypes.c_double(0); r=ll.microbench(w,0,ctypes.byref(o)); print(n, r, '%.4e'%o.value, 'per SM per clk %.1f' % (o.value/148/1.965e9))
"
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo rc=$?; tail -c 600 $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo rc=$?; tail -c 300 $OUT/bench_reference.err
