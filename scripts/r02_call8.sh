#!/usr/bin/env bash
# warp-level barriers, persistent Bluestein passes, dot_z / fold fusion: parity, racecheck, A/B against the CTA-barrier build
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02h.log 2>&1; tail -3 $OUT/pytest_gpu_r02h.log
for pass in 1 2; do
  timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02h_$pass.log 2>&1
  XLPROP_LIB=$PWD/build/libxlprop_nosw.so timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02h_nosw_$pass.log 2>&1
done
grep -- "---" $OUT/kern_r02h_1.log $OUT/kern_r02h_nosw_1.log $OUT/kern_r02h_2.log $OUT/kern_r02h_nosw_2.log
cat $OUT/kern_r02h_2.log
for m in grad vrsgrad cztgrad vcztgrad; do
    timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 512 $m 1 > $OUT/racecheck_${m}_r02h.log 2>&1
    tail -2 $OUT/racecheck_${m}_r02h.log
done
timeout 60 python scripts/gpu_probe.py --nosmoke --only2048 > $OUT/probe_r02h.log 2>&1; cat $OUT/probe_r02h.log
