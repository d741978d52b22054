#!/usr/bin/env bash
# round 2, multi-GPU record (run under `gpurun --gpus 8`): bench.py at 1/2/4/8 GPUs with the config-3/4/5 workloads under
# "extra", and the 16384^2 slab-decomposed RS parity check (point source vs analytic impulse response) at 2/4/8 GPUs.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_multi_n1.json 2> $OUT/bench_multi_n1.err
for n in 2 4 8; do
    timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 > $OUT/bench_multi_n$n.json 2> $OUT/bench_multi_n$n.err
done
for n in 1 2 4 8; do python - <<PY
import json
try:
    b = json.loads([l for l in open("$OUT/bench_multi_n$n.json") if l.startswith("{")][-1])
    print("N=$n value", round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), "extra", json.dumps(b["extra"])[:900])
except Exception as e:
    print("N=$n failed", e)
PY
done
for n in 2 4 8; do
    timeout 240 $TR --nproc-per-node $n --master-port $((29600 + n)) scripts/slab_check.py 16384 > $OUT/slab16384_n$n.json 2> $OUT/slab16384_n$n.err
    tail -c 600 $OUT/slab16384_n$n.json
done
timeout 200 python scripts/long_check.py 16384 > $OUT/long16384_n1.json 2> $OUT/long16384_n1.err; tail -c 500 $OUT/long16384_n1.json
for n in 2 8; do
    timeout 120 $TR --nproc-per-node $n --master-port $((29700 + n)) scripts/four_f_sharded.py --graph > $OUT/four_f_graph_n$n.json 2> $OUT/four_f_graph_n$n.err; tail -c 300 $OUT/four_f_graph_n$n.json
    timeout 120 $TR --nproc-per-node $n --master-port $((29800 + n)) scripts/four_f_sharded.py > $OUT/four_f_eager_n$n.json 2> $OUT/four_f_eager_n$n.err; tail -c 300 $OUT/four_f_eager_n$n.json
done
tail -3 $OUT/*.err | tail -40
