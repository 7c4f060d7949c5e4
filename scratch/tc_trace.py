import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, S, B = 20, 2, 8192
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, seed=1)
rng = np.random.default_rng(0)
node, edge, adj = synth_numpy(2048, N, rng)
nd, ed, ad = (torch.from_numpy(np.tile(t, (B // 2048, 1, 1))).cuda() for t in (node, edge, adj))
im, _, _ = v2v.pack_adjacency(ad)
q = torch.empty(B, N, 4, device="cuda")
lib = brain._lib
tr = torch.zeros(128, dtype=torch.int64, device="cuda")
nl = C.c_int32()
for _ in range(3):
    v2v._lib.check(lib.v2v_brain_tc_debug(brain._handle, nd.data_ptr(), ed.data_ptr(), im.data_ptr(), B, -2, q.data_ptr(), tr.data_ptr(), C.byref(nl), None))
torch.cuda.synchronize()
t = tr.cpu().numpy(); L = nl.value
print("tile start -> inputs loaded:", t[1] - t[0])
names = [f"stage{s}" for s in range(S)] + ["mlp1", "mlp2", "mlp3", "mlp4"]
prev = t[1]
for l in range(L):
    a, b_, c, d = t[2 + 4 * l: 6 + 4 * l]
    print(f"{names[l]:7s} split+sync {a - prev:6d} | issue {b_ - a:6d} | wait {c - b_:6d} | epilogue(+agg) {d - c:6d}")
    prev = d
print("tile total", prev - t[0], "cycles")
