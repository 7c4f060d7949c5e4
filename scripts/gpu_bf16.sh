#!/bin/bash
# bf16 tensor-core brain: tests + BASELINE configs[2] bench lines on one GPU.   usage: gpurun -- 'bash scripts/gpu_bf16.sh TAG'
TAG=${1:-r02}; O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_bf16.py -m gpu -q 2>&1 | tail -25
timeout 300 python bench.py --config c3 --steps 200 --warmup 10 > $O/bench_c3_weak_n1_$TAG.json 2> $O/bench_c3_weak_n1_$TAG.err; echo "c3 weak rc=$?"; tail -3 $O/bench_c3_weak_n1_$TAG.err
timeout 300 python bench.py --config c3 --scaling strong --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c3_strong_n1_$TAG.json 2> $O/bench_c3_strong_n1_$TAG.err; echo "c3 strong rc=$?"; tail -3 $O/bench_c3_strong_n1_$TAG.err
python - <<PY
import json
for f in ("$O/bench_c3_weak_n1_$TAG.json", "$O/bench_c3_strong_n1_$TAG.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "launches", d["gpu_launches"], "loss", d["final_loss"])
    except Exception as e:
        print(f, "ERR", e)
PY
