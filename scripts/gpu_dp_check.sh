#!/bin/bash
# One multi-GPU gpurun call: DP parity tests + bench at N GPUs (peer-memory exchange and NCCL backends).
TAG=${1:-r01}; NG=${2:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_dp_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_dp_$TAG.log
for be in peer nccl; do
V2V_DP_BACKEND=$be timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $NG --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_n${NG}_${be}_$TAG.json 2> $O/bench_n${NG}_${be}_$TAG.err; echo "bench $be rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_n${NG}_${be}_$TAG.json").read().strip().splitlines()[-1])
    print("$be", d["value"], d["ms_per_step"], d["e2e"]["value"] if d["e2e"] else None, d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["serialized"]["avg_launch_us"])
except Exception as e: print("parse fail", e)
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus $NG --config c3 --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_c3_weak_n${NG}_$TAG.json 2> $O/bench_c3_weak_n${NG}_$TAG.err; echo "c3 weak rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29545 \
  bench.py --gpus $NG --config c3 --scaling strong --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_c3_strong_n${NG}_$TAG.json 2> $O/bench_c3_strong_n${NG}_$TAG.err; echo "c3 strong rc=$?"
python - <<PY
import json
for f in ("bench_c3_weak_n${NG}", "bench_c3_strong_n${NG}"):
    try:
        d=json.loads(open("$O/%s_$TAG.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"] if d["e2e"] else None)
    except Exception as e: print(f, "parse fail", e)
PY
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_n1_$TAG.json 2> $O/bench_n1_$TAG.err; echo "bench1 rc=$?"; cat $O/bench_n1_$TAG.json | cut -c1-400
tail -3 $O/bench_n${NG}_peer_$TAG.err
