#!/bin/bash
# One multi-GPU gpurun call: scripts/sweep_dp.py at 1, 2, 4, 8 GPUs (as many as the box has) and the table.
# usage: gpurun --gpus 8 -- 'bash scripts/gpu_sweep_dp.sh TAG'
TAG=${1:-r02}; O=gpurun_out; mkdir -p $O
NGPU=$(nvidia-smi -L | wc -l)
: > $O/sweep_dp_$TAG.jsonl
for n in 1 2 4 8; do
  [ "$n" -le "$NGPU" ] || continue
  if [ "$n" -eq 1 ]; then
    timeout 400 python scripts/sweep_dp.py >> $O/sweep_dp_$TAG.jsonl 2> $O/sweep_dp_n${n}_$TAG.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      scripts/sweep_dp.py >> $O/sweep_dp_$TAG.jsonl 2> $O/sweep_dp_n${n}_$TAG.err
  fi
  echo "sweep_dp n=$n rc=$?"
done
python - <<PY | tee $O/sweep_dp_$TAG.txt
import json
rows = [json.loads(l) for l in open("$O/sweep_dp_$TAG.jsonl") if l.startswith("{")]
pts = sorted({(r["nodes"], r["graphs_per_gpu"], r["stages"], r["dtype"]) for r in rows}, key=lambda t: (t[3], t[0], t[1]))
gp = sorted({r["n_gpus"] for r in rows})
print("# data-parallel train step (fwd + Huber + bwd + NVLink gradient exchange + Keras-Adam), weak scaling: graphs per GPU fixed")
print("# M graphs/s in total (us per step; speed-up over 1 GPU); time = max over ranks, CUDA events, 200 steps after 20 warm-up")
print(f"{'nodes':>5} {'B/GPU':>6} {'stages':>6} {'dtype':>5} | " + " | ".join(f"{str(n) + ' GPU':>24}" for n in gp))
for p in pts:
    cells, base = [], None
    for n in gp:
        r = [x for x in rows if (x["nodes"], x["graphs_per_gpu"], x["stages"], x["dtype"]) == p and x["n_gpus"] == n]
        if not r:
            cells.append(f"{'-':>24}"); continue
        r = r[0]
        base = base or r["graphs_per_s"] / r["n_gpus"] * 1.0 if n == gp[0] else base
        sp = r["graphs_per_s"] / base if base else float("nan")
        cells.append(f"{r['graphs_per_s'] / 1e6:8.2f} ({r['us_per_step']:6.1f}; x{sp:4.2f})")
    print(f"{p[0]:>5} {p[1]:>6} {p[2]:>6} {p[3]:>5} | " + " | ".join(cells))
PY
