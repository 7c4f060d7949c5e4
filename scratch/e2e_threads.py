"""BS.train_dnn (packed fp32 host arrays, B=1024 x N=20) against the size of the host staging pool (V2V_HOST_THREADS)."""
import os, subprocess, sys
code = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, B = 20, 1024
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=B, data_parallel=False, seed=1)
rng = np.random.default_rng(0)
host = []
for i in range(8):
    node, edge, adj = synth_numpy(B, N, rng)
    host.append(({"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}, {"Decide_Output": rng.normal(0, 1, (B, N, 4)).astype(np.float32)}))
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for i in range(200):
        brain.train_dnn(host[i % 8][0], host[i % 8][1], B)
    best = min(best, (time.perf_counter() - t0) / 200 * 1e6)
print(f"threads {brain._lib.v2v_host_stage_threads()}: {best:.1f} us/call")
'''
for n in (2, 4, 6, 8, 10, 12, 15):
    env = dict(os.environ, V2V_HOST_THREADS=str(n))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
