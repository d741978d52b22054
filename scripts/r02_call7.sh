#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or four_f or fused or vs_oracle" > $OUT/pytest_gpu_r02g.log 2>&1; tail -2 $OUT/pytest_gpu_r02g.log
timeout 100 python scripts/four_f_profile.py 8 > $OUT/four_f_profile_b8.txt 2>&1; head -50 $OUT/four_f_profile_b8.txt
for b in 64 8; do
  timeout 100 python scripts/four_f_sharded.py --batch $b --graph > $OUT/four_f_1gpu_b${b}_graph_r02g.json 2> $OUT/ff1.err; tail -c 330 $OUT/four_f_1gpu_b${b}_graph_r02g.json | head -c 120; echo
done
timeout 90 python scripts/sharp_focus_table.py > $OUT/sharp_focus_r02g.json 2>&1; tail -c 300 $OUT/sharp_focus_r02g.json
timeout 60 python scripts/gpu_probe.py --nosmoke > $OUT/probe_r02g.log 2>&1; head -14 $OUT/probe_r02g.log
