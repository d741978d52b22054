#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -s > $OUT/pytest_gpu_r02d.log 2>&1; tail -3 $OUT/pytest_gpu_r02d.log; grep "d/dz\|four_f distance\|4f table\|sharp focus" $OUT/pytest_gpu_r02d.log
timeout 120 python scripts/kern_probe.py > $OUT/kern_r02d.log 2>&1; cat $OUT/kern_r02d.log
timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 128 vrsgrad 1 > $OUT/racecheck_vrsgrad_r02d.log 2>&1; tail -1 $OUT/racecheck_vrsgrad_r02d.log
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python scripts/prof_rs.py 100 grad 1 > $OUT/memcheck_grad_r02d.log 2>&1; tail -1 $OUT/memcheck_grad_r02d.log
