#!/bin/bash
OUT=gpurun_out/${1:-r3y}
mkdir -p $OUT
timeout 150 python -m pytest tests/test_slab_gpu.py -m gpu -q -x -k "_sort_levels or outside or cluster or wide" > $OUT/pytest_slab.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_slab.log
