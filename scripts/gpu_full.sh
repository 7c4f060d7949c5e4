#!/bin/bash
# One gpurun call: the whole GPU suite, smoke, both bench arms, the bf16 config, ncu launch lists and full captures.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_full.sh TAG'
TAG=${1:-r02}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi_$TAG.txt 2>&1; nproc >> $O/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke_$TAG.log
timeout 500 python bench.py > $O/bench_n1_$TAG.json 2> $O/bench_n1_$TAG.err; echo "bench rc=$?"; tail -2 $O/bench_n1_$TAG.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 300 python bench.py --config c3 --steps 200 --warmup 10 > $O/bench_c3_weak_n1_$TAG.json 2> $O/bench_c3_weak_n1_$TAG.err; echo "c3 weak rc=$?"
timeout 300 python bench.py --config c3 --scaling strong --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_c3_strong_n1_$TAG.json 2> $O/bench_c3_strong_n1_$TAG.err; echo "c3 strong rc=$?"
timeout 300 python bench.py --config c1 > $O/bench_c1_$TAG.json 2> $O/bench_c1_$TAG.err; echo "c1 rc=$?"; tail -2 $O/bench_c1_$TAG.err
timeout 300 python bench.py --config c4 > $O/bench_c4_$TAG.json 2> $O/bench_c4_$TAG.err; echo "c4 rc=$?"; tail -2 $O/bench_c4_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
   python bench.py --steps 20 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline > $O/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3_$TAG.csv \
   python bench.py --config c3 --steps 20 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline --no-e2e > $O/ncu_launches_c3_$TAG.log 2>&1; echo "ncu launches c3 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agg_mask -c 2 -o $O/agg_full_$TAG -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_agg_$TAG.log 2>&1; echo "ncu agg rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_brain -s 12 -c 1 -o $O/fused_full_$TAG -f \
   python bench.py --steps 6 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline --no-e2e > $O/ncu_fused_$TAG.log 2>&1; echo "ncu fused rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tt_kernel -s 12 -c 1 -o $O/tt_full_$TAG -f \
   python bench.py --config c3 --scaling strong --steps 6 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline --no-e2e > $O/ncu_tt_$TAG.log 2>&1; echo "ncu tt rc=$?"
python - <<PY
import json
for f in ("bench_n1", "bench_ref", "bench_c3_weak_n1", "bench_c3_strong_n1", "bench_c1", "bench_c4"):
    try:
        d = json.loads(open(f"$O/{f}_$TAG.json").read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "ms/step", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "launches", d.get("gpu_launches"))
        if d.get("roofline"): print("   roof frac", d["roofline"]["frac"], "dep", d["roofline"].get("frac_dependent"), "us", d["roofline"]["avg_launch_us"], d["roofline"]["serialized"]["avg_launch_us"])
        if d.get("points"): print("   ", json.dumps(d["points"]))
        if (d.get("e2e") or {}).get("reference_format"): print("   e2e ref format", d["e2e"]["reference_format"]["value"], d["e2e"]["reference_format"]["ms_per_call"])
    except Exception as e:
        print(f, "ERR", repr(e))
PY
ls -la $O | tail -12
# memcheck of the round-2 kernels on small cases (bf16 tcgen05 training kernel, per-slot fused kernel)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_brain.py -m gpu -q \
  -k "test_bf16_train_step and 7-1-50 or test_bf16_forward and 2-2-9 or test_fused_kernel_matches_layered_kernels and 4-3-5-True" \
  > $O/sanitizer_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_$TAG.log
