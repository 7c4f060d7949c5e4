#!/bin/bash
# Final round-2 evidence after the tensor-core backward of the fp32 fused kernel: GPU suite, smoke, both bench arms as the
# driver runs them, a steady-state bench line, the ncu launch list, memcheck of the new phases.
# usage: gpurun --timeout 600 -- 'bash scripts/gpu_final_r02.sh TAG'
TAG=${1:-r02b}; O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
timeout 60 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke_$TAG.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver_$TAG.json 2> $O/bench_driver_$TAG.err; echo "bench(driver flags) rc=$?"
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > $O/bench_n1_$TAG.json 2> $O/bench_n1_$TAG.err; echo "bench(300) rc=$?"
timeout 200 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
   python bench.py --steps 20 --warmup 3 --pool 4 --no-cpu-baseline --no-roofline > $O/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_brain.py -m gpu -q \
  -k "test_tensor_core_backward_matches_fp32_pipe and (4-3-64 or 9-1-50)" > $O/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck_$TAG.log
python - <<PY
import json
for f in ("bench_driver", "bench_n1", "bench_ref"):
    try:
        d = json.loads(open(f"$O/{f}_$TAG.json").read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "ms/step", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "launches", d.get("gpu_launches"))
        if d.get("roofline"): print("   roof frac", d["roofline"]["frac"], "dep", d["roofline"].get("frac_dependent"))
        if d.get("cpu_baseline"): print("   cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"))
    except Exception as e:
        print(f, "ERR", repr(e))
PY
