import os, sys, numpy as np, torch
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
brain.set_tensor_core(2)
for _ in range(3): brain.forward_device(nd, ed, in_mask=im, out=q)
torch.cuda.synchronize()
