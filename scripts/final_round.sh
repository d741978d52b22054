#!/usr/bin/env bash
# Final record of a tree in one short gpurun call (1 GPU): GPU suite, smoke, bench line, ncu launch list of the bench command.
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash scripts/final_round.sh r02w'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -2 $OUT/pytest_gpu_$TAG.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -1 $OUT/smoke_$TAG.log
timeout 300 python bench.py > $OUT/bench_${TAG}_final.json 2> $OUT/bench_${TAG}_final.err; tail -c 600 $OUT/bench_${TAG}_final.json; tail -2 $OUT/bench_${TAG}_final.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:xl_kernel -c 500 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > $OUT/bench_under_ncu.log 2>&1
tail -2 $OUT/launches_$TAG.csv | cut -c1-200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt
