"""Host-side breakdown of BS.train_dnn with numpy inputs (V2V_HOST_TRACE=1)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, B = 20, 1024
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=B, data_parallel=False, seed=1)
print("host threads", brain._lib.v2v_host_stage_threads())
rng = np.random.default_rng(0)
host = []
for i in range(8):
    node, edge, adj = synth_numpy(B, N, rng)
    host.append(({"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}, {"Decide_Output": rng.normal(0, 1, (B, N, 4)).astype(np.float32)}))
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(200):
        brain.train_dnn(host[i % 8][0], host[i % 8][1], B)
    dt = time.perf_counter() - t0
    print(f"packed fp32: {dt / 200 * 1e6:.1f} us/call", flush=True)
# reference form: per-slot fp64 dict
ref = []
for i in range(4):
    node, edge, adj = synth_numpy(B, N, rng)
    x = {"Adjacency_Matrix": adj.astype(np.float64)}
    for k in range(N):
        x[f"D{k+1}_Node_Input"] = node[:, k].astype(np.float64); x[f"D{k+1}_Edge_Input"] = edge[:, k].astype(np.float64)
        x[f"D{k+1}_Neighbor_Input"] = np.zeros((B, 16))
    y = {f"D{k+1}_Decide_Output": rng.normal(0, 1, (B, 4)) for k in range(N)}
    ref.append((x, y))
for rep in range(2):
    t0 = time.perf_counter()
    for i in range(200):
        brain.train_dnn(ref[i % 4][0], ref[i % 4][1], B)
    dt = time.perf_counter() - t0
    print(f"per-slot fp64 dict: {dt / 200 * 1e6:.1f} us/call", flush=True)
t0 = time.perf_counter()
for i in range(200):
    brain.predict(host[i % 8][0])
print(f"predict packed fp32: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us/call")
