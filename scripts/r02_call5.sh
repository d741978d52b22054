#!/usr/bin/env bash
# round 2: fused 4f elements + transposed CZT factor table: full GPU suite, per-kernel timings, 4f step fused vs unfused
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -s > $OUT/pytest_gpu_r02e.log 2>&1; tail -3 $OUT/pytest_gpu_r02e.log; grep "4f fused\|4f table" $OUT/pytest_gpu_r02e.log
timeout 150 python scripts/kern_probe.py > $OUT/kern_r02e.log 2>&1; grep -A8 "CZT\|highNA" $OUT/kern_r02e.log
for b in 64 8; do
  timeout 100 python scripts/four_f_sharded.py --batch $b > $OUT/four_f_b${b}_fused_r02e.json 2> $OUT/four_f_fused.err; tail -c 400 $OUT/four_f_b${b}_fused_r02e.json
  timeout 100 python scripts/four_f_sharded.py --batch $b --unfused > $OUT/four_f_b${b}_unfused_r02e.json 2>> $OUT/four_f_fused.err; tail -c 400 $OUT/four_f_b${b}_unfused_r02e.json
done
timeout 100 python scripts/four_f_sharded.py --batch 8 --graph > $OUT/four_f_b8_fused_graph_r02e.json 2>> $OUT/four_f_fused.err; tail -c 400 $OUT/four_f_b8_fused_graph_r02e.json
tail -5 $OUT/four_f_fused.err
