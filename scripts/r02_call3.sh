#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or rs_ or vrs_ or four_f or slab or split" > $OUT/parity_r02c.log 2>&1; tail -3 $OUT/parity_r02c.log
timeout 120 python scripts/kern_probe.py > $OUT/kern_r02c_async.log 2>&1
XL_ROWS_ASYNC=0 timeout 120 python scripts/kern_probe.py > $OUT/kern_r02c_rowsclassic.log 2>&1
for f in kern_r02c_async kern_r02c_rowsclassic; do echo "== $f"; cat $OUT/$f.log; done
