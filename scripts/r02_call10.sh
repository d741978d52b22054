#!/usr/bin/env bash
# overlapped twiddle fetch, K4 input staged with cp.async: parity, racecheck, timings, bench value; fused AdamW A/B on cfg 4
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02j.log 2>&1; tail -3 $OUT/pytest_gpu_r02j.log
timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02j.log 2>&1; cat $OUT/kern_r02j.log
for m in grad vrsgrad cztgrad; do
    timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 512 $m 1 > $OUT/racecheck_${m}_r02j.log 2>&1
    tail -2 $OUT/racecheck_${m}_r02j.log
done
timeout 100 compute-sanitizer --tool memcheck --print-limit 5 python scripts/prof_rs.py 512 grad 1 > $OUT/memcheck_grad_r02j.log 2>&1; tail -2 $OUT/memcheck_grad_r02j.log
timeout 60 python scripts/gpu_probe.py --nosmoke --only2048 > $OUT/probe_r02j.log 2>&1; cat $OUT/probe_r02j.log
timeout 200 python bench.py --no-cpu-baseline --no-extra > $OUT/bench_r02j.json 2> $OUT/bench_r02j.err; head -c 900 $OUT/bench_r02j.json; echo
for fa in "" "--fused-adam"; do
  timeout 100 python scripts/four_f_sharded.py --batch 8 --graph $fa > $OUT/four_f_b8_r02j$fa.json 2>> $OUT/ff1.err; echo "b8 graph $fa: $(grep -o '"ms_per_step": [0-9.]*' $OUT/four_f_b8_r02j$fa.json)"
done
timeout 100 python scripts/four_f_sharded.py --batch 64 --graph --fused-adam > $OUT/four_f_b64_r02j.json 2>> $OUT/ff1.err; echo "b64 graph fused-adam: $(grep -o '"ms_per_step": [0-9.]*' $OUT/four_f_b64_r02j.json)"
