import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N = 20
for S in (2, 3):
  for B in (1024, 2048, 4096, 8192, 32768):
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, seed=1)
    rng = np.random.default_rng(0)
    node, edge, adj = synth_numpy(min(B, 2048), N, rng)
    rep = B // node.shape[0]
    nd, ed, ad = (torch.from_numpy(np.tile(t, (rep, 1, 1))).cuda() for t in (node, edge, adj))
    im, _, _ = v2v.pack_adjacency(ad)
    q = torch.empty(B, N, 4, device="cuda")
    res = {}
    for mode in (0, 2):
        brain.set_tensor_core(mode)
        for _ in range(5): brain.forward_device(nd, ed, in_mask=im, out=q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): brain.forward_device(nd, ed, in_mask=im, out=q)
        e1.record(); torch.cuda.synchronize()
        res[mode] = 1e3 * e0.elapsed_time(e1) / 50
    print(f"S={S} B={B}: FP32-pipe {res[0]:.1f} us, tcgen05 3xTF32 {res[2]:.1f} us  ({res[0]/res[2]:.2f}x)  {B/res[2]:.1f} graphs/us", flush=True)
