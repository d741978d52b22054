#!/usr/bin/env bash
# quick GPU iteration: parity gate on the RS family + per-kernel timings
set -u
OUT=gpurun_out
TAG=${1:-q}
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_2048.py -q -x -k "${2:-golden or rs_ or vrs_ or four_f}" > $OUT/parity_$TAG.log 2>&1; tail -2 $OUT/parity_$TAG.log
timeout 120 python scripts/kern_probe.py > $OUT/kern_$TAG.log 2>&1; cat $OUT/kern_$TAG.log
