#!/bin/bash
# 8-GPU box: end-to-end train_dnn throughput against the number of staging workers per rank (default: cores / ranks - 1 = 3)
N=$(nvidia-smi -L | wc -l)
for t in 3 2 1 5; do
  V2V_HOST_THREADS=$t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$N GPUs, workers $t: e2e', round(d['e2e']['value']/1e6,2), 'M graphs/s  value', round(d['value']/1e6,2))"
done
lscpu | grep -E "Thread|Core|Socket|Model name|NUMA" 
