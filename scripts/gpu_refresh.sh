#!/bin/bash
# One single-GPU gpurun call after a kernel change: GPU suite, smoke, both bench arms, ncu launch list + full capture of the fused kernel.
TAG=r02n; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 500 python bench.py > $O/bench_n1_$TAG.json 2> $O/bench_n1_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 20 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline > $O/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_brain -s 12 -c 1 -o $O/fused_full_$TAG -f python bench.py --steps 6 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline --no-e2e > $O/ncu_fused_$TAG.log 2>&1; echo "ncu fused rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_n1_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["frac_dependent"], "cpu", d["cpu_baseline"]["value"], "predict", d["predict"]["tcgen05_3xtf32_us"])
PY
