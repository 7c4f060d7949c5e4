"""Where does BS.train_dnn / BS.predict spend its host time at the reference's shape (N=4, per-slot, fp64 dict, Kronecker A)?"""
import os, sys, time, cProfile, pstats, io
os.environ["V2V_HOST_TRACE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, F, B = 4, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 256
brain = v2v.BS(N, 3, 1, 16, 1, 4, data_parallel=False, seed=1, per_slot=True)
rng = np.random.default_rng(0)
node, edge, adj = (t.astype(np.float64) for t in synth_numpy(B, N, rng))
A = np.stack([np.kron(a, np.eye(F)) for a in adj])
x = {"Adjacency_Matrix": A}
for k in range(N):
    x[f"D{k+1}_Node_Input"] = node[:, k]; x[f"D{k+1}_Edge_Input"] = edge[:, k]; x[f"D{k+1}_Neighbor_Input"] = np.zeros((B, F))
y = {f"D{k+1}_Decide_Output": rng.normal(size=(B, 4)) for k in range(N)}
for _ in range(50): brain.predict(x); brain.train_dnn(x, y, B)
torch.cuda.synchronize()
for name, fn in (("train_dnn", lambda: brain.train_dnn(x, y, B)), ("predict", lambda: brain.predict(x))):
    t0 = time.perf_counter()
    for _ in range(300): fn()
    dt = (time.perf_counter() - t0) / 300 * 1e6
    pr = cProfile.Profile(); pr.enable()
    for _ in range(300): fn()
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12)
    print(f"==== {name}: {dt:.1f} us per call (unprofiled)"); print("\n".join(s.getvalue().splitlines()[4:24]))
