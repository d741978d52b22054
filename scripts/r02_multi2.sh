#!/usr/bin/env bash
# round 2, final tree: bench.py (with the config-3/4/5 workloads under "extra") at 2 and 8 GPUs of one box
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi_multi2.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 2; do
    timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 > $OUT/bench_multi2_n$n.json 2> $OUT/bench_multi2_n$n.err
    python - <<PY
import json
try:
    b = json.loads([l for l in open("$OUT/bench_multi2_n$n.json") if l.startswith("{")][-1])
    print("N=$n value", round(b["value"], 1), "e2e", round(b["e2e"]["value"], 1), "extra", json.dumps(b["extra"])[:1200])
except Exception as e:
    print("N=$n failed", e)
PY
done
tail -3 $OUT/bench_multi2_n8.err
