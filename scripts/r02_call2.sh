#!/usr/bin/env bash
# Round 2, call 2: bulk-asynchronous rs_cols / rs_rows_fwd against the round-1 library (parity gate, per-kernel timings,
# stagger knob, racecheck of the mbarrier pipeline).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_2048.py -q -x -k "golden or rs_ or vrs_ or four_f" > $OUT/parity_r02b.log 2>&1; tail -3 $OUT/parity_r02b.log
XLPROP_LIB=$PWD/build/libxlprop_r01.so timeout 120 python scripts/kern_probe.py > $OUT/kern_r01lib.log 2>&1
timeout 120 python scripts/kern_probe.py > $OUT/kern_r02b_s0.log 2>&1
XL_STAGGER_NS=3000 timeout 120 python scripts/kern_probe.py > $OUT/kern_r02b_s3000.log 2>&1
XL_STAGGER_NS=7000 timeout 120 python scripts/kern_probe.py > $OUT/kern_r02b_s7000.log 2>&1
for f in kern_r01lib kern_r02b_s0 kern_r02b_s3000 kern_r02b_s7000; do echo "== $f"; head -22 $OUT/$f.log; done
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 128 grad 1 > $OUT/racecheck_grad_r02b.log 2>&1; tail -1 $OUT/racecheck_grad_r02b.log
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python scripts/prof_rs.py 128 vrsgrad 1 > $OUT/memcheck_vrsgrad_r02b.log 2>&1; tail -1 $OUT/memcheck_vrsgrad_r02b.log
