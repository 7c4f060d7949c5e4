#!/usr/bin/env python
"""BASELINE configs[4], the multi-GPU part: the data-parallel train step (fwd + Huber + bwd + NVLink gradient exchange +
Keras-Adam) over graph size N and per-GPU batch B at WORLD_SIZE GPUs, weak scaling (B graphs on every GPU).  One process
group per launch, every point inside it; rank 0 prints one JSON line per point.  Time = max over ranks, CUDA events.

    python scripts/sweep_dp.py                                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 scripts/sweep_dp.py

scripts/gpu_sweep_dp.sh runs it at 1, 2, 4, 8 GPUs and builds the table (profiles/sweep_dp_rNN.txt)."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_numpy, SEED      # noqa: E402

POINTS = [  # (nodes, graphs per GPU, stages, dtype)
    (8, 1024, 2, "f32"), (8, 4096, 2, "f32"),
    (20, 256, 2, "f32"), (20, 1024, 2, "f32"), (20, 4096, 2, "f32"),
    (32, 1024, 2, "f32"),
    (20, 1024, 3, "bf16"), (20, 4096, 3, "bf16"),
]


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    v2v = importlib.import_module("globecom2020-resourceallocationgnn_b200")
    lib = v2v.load_library()
    ptr = v2v._lib.ptr
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    steps, warm = 200, 20
    for N, B, S, dtype in POINTS:
        brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=(world > 1), seed=SEED, dtype=dtype)
        brain.update_target_model()
        if world > 1:
            for w in (0, 1):
                dist.broadcast(brain._views[w], src=0)
        rng = np.random.default_rng(SEED + rank)
        pool = []
        for i in range(4):
            node, edge, adj = synth_numpy(B, N, rng)
            nd, ed, ad = (torch.from_numpy(t).to(dev) for t in (node, edge, adj))
            im, om, _ = v2v.pack_adjacency(ad)
            p = brain.forward_device(nd, ed, in_mask=im)
            pn = brain.forward_device(nd, ed, in_mask=im, target=True)
            act = torch.from_numpy(rng.integers(0, 4, (B, N)).astype(np.int32)).to(dev)
            rew = torch.from_numpy(rng.normal(10.0, 3.0, B).astype(np.float32)).to(dev)
            y = torch.empty_like(p)
            v2v._lib.check(lib.v2v_td_target(ptr(p), ptr(pn), ptr(act), ptr(rew), 0.5, ptr(y), B, N, 4, v2v._lib.current_stream()))
            pool.append((nd, ed, im, om, y))
        head_loss = torch.zeros(N, device=dev)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for i in range(warm):
            brain.train_step_device(*pool[i % 4][:4], None, pool[i % 4][4], head_loss=head_loss)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            brain.train_step_device(*pool[i % 4][:4], None, pool[i % 4][4], head_loss=head_loss)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        loss = float(head_loss.sum().item())
        assert np.isfinite(loss)
        if rank == 0:
            print(json.dumps({"n_gpus": world, "nodes": N, "graphs_per_gpu": B, "stages": S, "dtype": dtype,
                              "us_per_step": 1e3 * float(ms.item()) / steps,
                              "graphs_per_s": world * B * steps / (float(ms.item()) * 1e-3)}), flush=True)
        del brain, pool
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
