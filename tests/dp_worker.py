"""Run under torchrun with >= 2 GPUs: the data-parallel train step with (a) the fused NVLink peer exchange + Adam kernel
and (b) NCCL all-reduce + Adam must both reproduce the single-process full-batch step (tests/test_gpu_dp.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v                      # noqa: E402
from oracle import v2v_oracle as O               # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    N, S, Bl = 20, 2, 96
    rng = np.random.default_rng(5)
    node, edge, adj, _ = O.synth_batch(Bl * world, N, rng)
    node, edge, adj = node.astype(np.float32), edge.astype(np.float32), adj.astype(np.float32)
    y = rng.normal(0, 1, (Bl * world, N, 4)).astype(np.float32)
    d = O.BrainDims(N, stages=S, per_slot=False)
    p0 = O.flatten_params(O.init_params(d, np.random.default_rng(9), bias_scale=0.05)).astype(np.float32)
    lo, hi = rank * Bl, (rank + 1) * Bl
    dev = lambda a: torch.from_numpy(a).cuda()
    results = {}
    for backend in ("peer", "nccl"):
        os.environ["V2V_DP_BACKEND"] = backend
        brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=Bl * world, data_parallel=True, seed=1)
        assert (brain._comm is not None) == (backend == "peer")
        brain.set_flat_params(p0, 0)
        im, om, _ = v2v.pack_adjacency(dev(adj[lo:hi]))
        losses = []
        for _ in range(3):
            hl = brain.train_step_device(dev(node[lo:hi]), dev(edge[lo:hi]), im, om, None, dev(y[lo:hi]))
            losses.append(hl.clone())
        if brain._comm is not None:
            v2v._lib.check(brain._lib.v2v_comm_check(brain._comm, v2v._lib.current_stream()))
        torch.cuda.synchronize()
        results[backend] = (brain.get_flat_params(0), brain.get_flat_params(2), torch.stack(losses).cpu().numpy())
        # a 4th step through the reference-facing host call (numpy in, per-head losses out) on this rank's rows
        hist = brain.train_dnn({"Node_Input": node[lo:hi], "Edge_Input": edge[lo:hi], "Adjacency_Matrix": adj[lo:hi]},
                               {"Decide_Output": y[lo:hi]}, Bl)
        results[backend + "_host"] = (brain.get_flat_params(0), hist.history["loss"][0])
        # every rank must hold bit-identical parameters
        mine = torch.from_numpy(results[backend][0]).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), f"{backend}: replicas diverged"
        del brain
    # single-process full-batch reference on rank 0's GPU (every rank computes it: cheap)
    os.environ["V2V_DP_BACKEND"] = "nccl"
    single = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=Bl * world, data_parallel=False, seed=1)
    single.set_flat_params(p0, 0)
    im, om, _ = v2v.pack_adjacency(dev(adj))
    sl = []
    for _ in range(3):
        sl.append(single.train_step_device(dev(node), dev(edge), im, om, None, dev(y)).clone())
    ps, gs = single.get_flat_params(0), single.get_flat_params(2)
    sl = torch.stack(sl).cpu().numpy()
    for backend in ("peer", "nccl"):
        p, g, l = results[backend]
        gerr = np.abs(g - gs).max() / np.abs(gs).max()
        assert gerr < 5e-5, (backend, gerr)
        assert np.abs(p - ps).max() < 2e-4 and np.quantile(np.abs(p - ps), 0.9) < 2e-6, backend
    hs = single.train_dnn({"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}, {"Decide_Output": y}, Bl * world)
    p4 = single.get_flat_params(0)
    for backend in ("peer", "nccl"):
        ph, lh = results[backend + "_host"]
        assert np.abs(ph - p4).max() < 3e-4 and np.quantile(np.abs(ph - p4), 0.9) < 3e-6, backend
        if backend == "peer":                 # global-mean loss on every rank
            assert abs(lh - hs.history["loss"][0]) < 1e-4 * abs(hs.history["loss"][0]), (lh, hs.history["loss"][0])
    # the peer path carries the per-head losses through the same exchange: global mean on every rank
    assert np.abs(results["peer"][2] - sl).max() < 1e-4 * np.abs(sl).max()
    # the bf16 tensor-core configuration (3 stages) through the same fused exchange: ranks' shards vs the full batch
    os.environ["V2V_DP_BACKEND"] = "peer"
    d3 = O.BrainDims(N, stages=3, per_slot=False)
    p3 = O.flatten_params(O.init_params(d3, np.random.default_rng(11), bias_scale=0.05)).astype(np.float32)
    bd = v2v.BS(N, 3, 1, 16, 1, 4, stages=3, per_slot=False, max_batch=Bl * world, data_parallel=True, seed=1, dtype="bf16")
    bs = v2v.BS(N, 3, 1, 16, 1, 4, stages=3, per_slot=False, max_batch=Bl * world, data_parallel=False, seed=1, dtype="bf16")
    bd.set_flat_params(p3, 0); bs.set_flat_params(p3, 0)
    im_l, om_l, _ = v2v.pack_adjacency(dev(adj[lo:hi]))
    q_full = bs.forward_device(dev(node), dev(edge), in_mask=im)
    y3 = (q_full + 0.4).contiguous()
    for _ in range(2):
        l_dp = bd.train_step_device(dev(node[lo:hi]), dev(edge[lo:hi]), im_l, om_l, None, y3[lo:hi].contiguous())
        l_sp = bs.train_step_device(dev(node), dev(edge), im, om, None, y3)
    v2v._lib.check(bd._lib.v2v_comm_check(bd._comm, v2v._lib.current_stream()))
    g_dp, g_sp = bd.get_flat_params(2), bs.get_flat_params(2)
    assert np.abs(g_dp - g_sp).max() <= 2e-3 * np.abs(g_sp).max(), np.abs(g_dp - g_sp).max() / np.abs(g_sp).max()
    assert np.abs(l_dp.cpu().numpy() - l_sp.cpu().numpy()).max() <= 1e-4 * np.abs(l_sp.cpu().numpy()).max()
    mine = torch.from_numpy(bd.get_flat_params(0)).cuda()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), "bf16: replicas diverged"
    dist.barrier()
    if rank == 0:
        print("DP_WORKER_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
