#!/bin/bash
# One multi-GPU gpurun call: the default bench line (BASELINE configs[1], weak scaling, with e2e) at 1, 2, 4, 8 GPUs.
# usage: gpurun --gpus 8 -- 'bash scripts/gpu_scale_c2.sh'   (scripts/gpu_scale.sh adds the configs[2] weak / strong lines)
TAG=${1:-r02m}; O=gpurun_out; mkdir -p $O
for n in 1 2 4 8; do
  if [ "$n" -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 300 --warmup 20 --no-cpu-baseline > $O/scale_c2_n${n}_$TAG.json 2> $O/scale_c2_n${n}_$TAG.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n --steps 300 --warmup 20 --no-cpu-baseline > $O/scale_c2_n${n}_$TAG.json 2> $O/scale_c2_n${n}_$TAG.err
  fi
  echo "n=$n rc=$?"
done
python - <<PY
import json
base=None
for n in (1,2,4,8):
    d=json.loads(open(f"$O/scale_c2_n{n}_$TAG.json").read().strip().splitlines()[-1])
    base = base or d["value"]
    print(n, round(d["value"]/1e6,2), round(1e3*d["ms_per_step"],1), "x%.2f" % (d["value"]/base), "e2e", round(d["e2e"]["value"]/1e6,2), d["e2e"].get("host_threads"))
PY
