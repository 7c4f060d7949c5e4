#!/bin/bash
# quick single-GPU check: GPU tests + bench
TAG=${1:-q}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu_$TAG.log
timeout 400 python bench.py ${BENCH_ARGS:---steps 300 --warmup 20} > $O/bench_n1_$TAG.json 2> $O/bench_n1_$TAG.err; echo "bench rc=$?"; tail -3 $O/bench_n1_$TAG.err
python - <<PY
import json
d=json.loads(open("$O/bench_n1_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], "\nroof", d["roofline"]["frac"], d["roofline"]["avg_launch_us"], d["roofline"]["serialized"]["avg_launch_us"], "cpu", d["cpu_baseline"])
PY
