"""Reference shape (C1/C4): N=4, per-slot weights, 3 stages, B=256/512, reference-format fp64 dict inputs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, F = 4, 16
for per_slot in (True, False):
  for B in (1, 256, 512):
    brain = v2v.BS(N, 3, 1, 16, 1, 4, data_parallel=False, seed=1, per_slot=per_slot)
    rng = np.random.default_rng(0)
    node, edge, adj = (t.astype(np.float64) for t in synth_numpy(B, N, rng))
    A = np.stack([np.kron(a, np.eye(F)) for a in adj])
    x = {"Adjacency_Matrix": A}
    for k in range(N):
        x[f"D{k+1}_Node_Input"] = node[:, k]; x[f"D{k+1}_Edge_Input"] = edge[:, k]; x[f"D{k+1}_Neighbor_Input"] = np.zeros((B, F))
    y = {f"D{k+1}_Decide_Output": rng.normal(size=(B, 4)) for k in range(N)}
    for _ in range(20): brain.predict(x); brain.train_dnn(x, y, B)
    torch.cuda.synchronize(); lc0 = brain._lib.v2v_launch_count()
    t0 = time.perf_counter()
    for _ in range(200): brain.predict(x)
    tp = (time.perf_counter() - t0) / 200 * 1e6; lc1 = brain._lib.v2v_launch_count()
    t0 = time.perf_counter()
    for _ in range(200): brain.train_dnn(x, y, B)
    tt = (time.perf_counter() - t0) / 200 * 1e6; lc2 = brain._lib.v2v_launch_count()
    print(f"per_slot={per_slot} B={B}: predict {tp:.0f} us ({(lc1-lc0)/200:.0f} launches), train_dnn {tt:.0f} us ({(lc2-lc1)/200:.0f} launches)", flush=True)
