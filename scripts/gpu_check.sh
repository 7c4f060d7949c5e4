#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, ncu launch list + full captures of the two top kernels.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi_$TAG.txt 2>&1
nproc >> $O/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
tail -5 $O/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$TAG.log
timeout 400 python bench.py > $O/bench_n1_$TAG.json 2> $O/bench_n1_$TAG.err; echo "bench rc=$?"; cat $O/bench_n1_$TAG.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"; cat $O/bench_ref_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_mask -c 2 -o $O/agg_full_$TAG -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_agg_$TAG.log 2>&1; echo "ncu agg rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_brain -s 200 -c 1 -o $O/fused_full_$TAG -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_fused_$TAG.log 2>&1; echo "ncu fused rc=$?"
ls -la $O | tail -20
