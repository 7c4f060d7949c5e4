"""Per-call wall time of BS.train_dnn in the order bench.py runs it (device steps first, 3 warm-up calls, 20 timed)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, B, CH = 20, 1024, 4
brain = v2v.BS(N, 3, 1, 16, 1, CH, stages=2, per_slot=False, max_batch=B, data_parallel=False, seed=1)
rng = np.random.default_rng(0)
host = []
for i in range(8):
    node, edge, adj = synth_numpy(B, N, rng)
    host.append(({"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}, {"Decide_Output": rng.normal(0, 1, (B, N, CH)).astype(np.float32)}))
for trial in range(3):
    time.sleep(0.5)
    ts = []
    for i in range(3 + 40):
        t0 = time.perf_counter()
        brain.train_dnn(host[i % 8][0], host[i % 8][1], B)
        ts.append((time.perf_counter() - t0) * 1e6)
    print("trial", trial, "warm", [f"{t:.0f}" for t in ts[:3]], "timed", [f"{t:.0f}" for t in ts[3:]], f"mean20 {np.mean(ts[3:23]):.1f} mean next20 {np.mean(ts[23:]):.1f}", flush=True)
