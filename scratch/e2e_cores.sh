#!/bin/bash
# Emulates the 8-GPU box's core budget (4 cores per rank) on a 2-GPU box: 2 ranks pinned to 8 cores, staging workers 1..3
for t in 3 2 1; do
  V2V_HOST_THREADS=$t taskset -c 0-7 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('8 cores, workers $t: e2e', round(d['e2e']['value']/1e6,2), 'M graphs/s  value', round(d['value']/1e6,2))"
done
V2V_HOST_THREADS=3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('all cores, workers 3: e2e', round(d['e2e']['value']/1e6,2))"
