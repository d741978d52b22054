#!/usr/bin/env bash
# 2 GPUs: cfg-4 step variants (eager / graphs + eager all-reduce / one graph with NCCL), GPU suite on the new ABI
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02f.log 2>&1; tail -3 $OUT/pytest_gpu_r02f.log
for b in 64 8; do
  timeout 100 python scripts/four_f_sharded.py --batch $b --graph > $OUT/four_f_1gpu_b${b}_graph_r02f.json 2> $OUT/ff1.err; tail -c 250 $OUT/four_f_1gpu_b${b}_graph_r02f.json; echo
done
timeout 100 python scripts/four_f_sharded.py --batch 8 > $OUT/four_f_1gpu_b8_eager_r02f.json 2>> $OUT/ff1.err; tail -c 250 $OUT/four_f_1gpu_b8_eager_r02f.json; echo
for mode in "" "--graph" "--graph-nccl"; do
  timeout 90 $TR --nproc-per-node 2 --master-port 29611 scripts/four_f_sharded.py --batch 16 $mode > $OUT/four_f_2gpu_b16_${mode#--}_r02f.json 2> $OUT/ff2_${mode#--}.err
  echo "2 GPUs batch 16 mode '$mode': rc=$? $(tail -c 250 $OUT/four_f_2gpu_b16_${mode#--}_r02f.json)"
done
tail -3 $OUT/ff2_graph-nccl.err
