#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02n.log 2>&1; tail -3 $OUT/pytest_gpu_r02n.log
timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02n.log 2>&1; sed -n '/CZT fwd/,$p' $OUT/kern_r02n.log
timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 512 vcztgrad 1 > $OUT/racecheck_vcztgrad_r02n.log 2>&1; tail -2 $OUT/racecheck_vcztgrad_r02n.log
