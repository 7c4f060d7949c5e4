"""Step time of the fused fp32 kernel (configs[1] shape) in both contraction modes, B = 1024 and 8192."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
lib = v2v.load_library()
out = []
for B in (1024, 8192):
    rng = np.random.default_rng(0)
    brain = v2v.BS(20, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=B, data_parallel=False, seed=9)
    node, edge, adj = synth_numpy(B, 20, rng)
    nd, ed, ad = (torch.from_numpy(t.astype(np.float32)).cuda() for t in (node, edge, adj))
    im, om, _ = v2v.pack_adjacency(ad)
    y = brain.forward_device(nd, ed, in_mask=im) + 1
    for mode in (0, 1):
        lib.v2v_fused_set_mma(mode)
        for _ in range(10): brain.train_step_device(nd, ed, im, om, None, y)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100): brain.train_step_device(nd, ed, im, om, None, y)
            e1.record(); torch.cuda.synchronize()
            best = min(best, 10 * e0.elapsed_time(e1))
        out.append(f"B={B} mode{mode} {best:.1f}us")
print(sys.argv[1] if len(sys.argv) > 1 else "", " | ".join(out), flush=True)
