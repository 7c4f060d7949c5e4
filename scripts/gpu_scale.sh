#!/bin/bash
# One multi-GPU gpurun call: scaling lines at N = 1, 2, 4, 8 (as many as the box has) for the default config (BASELINE
# configs[1], weak) and for configs[2] (bf16, 3 stages; weak and strong).   usage: gpurun --gpus 8 -- 'bash scripts/gpu_scale.sh TAG'
TAG=${1:-r02}; O=gpurun_out; mkdir -p $O
NGPU=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $O/topo_$TAG.txt 2>&1
run() {   # name n args...
  local name=$1 n=$2; shift 2
  if [ "$n" -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $O/${name}_n${n}_$TAG.json 2> $O/${name}_n${n}_$TAG.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      bench.py --gpus $n "$@" > $O/${name}_n${n}_$TAG.json 2> $O/${name}_n${n}_$TAG.err
  fi
  echo "$name n=$n rc=$?"
}
for n in 1 2 4 8; do
  [ "$n" -le "$NGPU" ] || continue
  run scale_c2 $n --steps 300 --warmup 20 --no-cpu-baseline
  run scale_c3_weak $n --config c3 --steps 300 --warmup 20 --no-cpu-baseline
  run scale_c3_strong $n --config c3 --scaling strong --steps 300 --warmup 20 --no-cpu-baseline
done
python - <<PY
import json, glob
for name in ("scale_c2", "scale_c3_weak", "scale_c3_strong"):
    base = None
    for n in (1, 2, 4, 8):
        try:
            d = json.loads(open(f"$O/{name}_n{n}_$TAG.json").read().strip().splitlines()[-1])
        except Exception as e:
            continue
        base = base or d["value"]
        e2e = d["e2e"]["value"] if d.get("e2e") else float("nan")
        print(f"{name:16s} n={n} value {d['value']/1e6:8.2f} M graphs/s  {1e3*d['ms_per_step']:7.1f} us/step  x{d['value']/base:5.2f} (eff {d['value']/base/n:4.2f})  e2e {e2e/1e6:7.2f} M  loss {d.get('final_loss')}")
PY
