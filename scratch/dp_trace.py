"""torchrun --nproc-per-node 2 scratch/dp_trace.py : phase trace of the fused gradient exchange kernel inside real train steps."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
N, B = 20, 1024
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=B, data_parallel=True, seed=1)
lib = brain._lib
rng = np.random.default_rng(rank)
node, edge, adj = synth_numpy(B, N, rng)
nd, ed, ad = (torch.from_numpy(t).to(dev) for t in (node, edge, adj))
im, om, _ = v2v.pack_adjacency(ad)
y = torch.randn(B, N, 4, device=dev)
nch = lib.v2v_comm_num_chunks(brain._comm)
tr = torch.zeros(nch * 6, dtype=torch.int64, device=dev)
for mode in ("steps", "comm_only"):
    for it in range(60):
        if it == 50:
            lib.v2v_comm_set_trace(brain._comm, tr.data_ptr())
        if it == 10:
            torch.cuda.synchronize(); dist.barrier(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        brain.train_step_device(nd, ed, im, om, None, y)
    e1.record(); torch.cuda.synchronize()
    lib.v2v_comm_set_trace(brain._comm, None)
    t = tr.cpu().numpy().reshape(nch, 6).astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t[:, :5] - t0) / 1e3
    print(f"rank {rank} {mode}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us/step; trace us (min/median/max over chunks): "
          + " | ".join(f"{nm} {rel[:, k].min():.1f}/{np.median(rel[:, k]):.1f}/{rel[:, k].max():.1f}"
                       for k, nm in enumerate(("entry", "producer_done", "pushed", "arrived", "done"))), flush=True)
    break
dist.barrier(); dist.destroy_process_group()
