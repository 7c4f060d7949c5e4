"""Per-step clock trace of the bf16 tensor-core training kernel (CTA 0, first two tiles): where does a tile's time go?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N, S, B = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 3, int(sys.argv[2]) if len(sys.argv) > 2 else 8192
lib = v2v.load_library()
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, seed=1, dtype="bf16")
rng = np.random.default_rng(0)
node, edge, adj = synth_numpy(min(B, 1024), N, rng)
rep = -(-B // node.shape[0])
nd, ed, ad = (torch.from_numpy(np.tile(t, (rep, 1, 1))[:B]).cuda() for t in (node, edge, adj))
im, om, _ = v2v.pack_adjacency(ad)
y = brain.forward_device(nd, ed, in_mask=im) + 0.3
for _ in range(3):
    brain.train_step_device(nd, ed, im, om, None, y)
buf = torch.zeros(2 * 24 * 8, dtype=torch.int64, device="cuda")
v2v._lib.check(lib.v2v_tt_set_trace(buf.data_ptr()))
brain.train_step_device(nd, ed, im, om, None, y)
torch.cuda.synchronize()
v2v._lib.check(lib.v2v_tt_set_trace(None))
t = buf.cpu().numpy().reshape(2, 24, 8)
for tile in range(2):
    t0 = t[tile, 23, 6]
    print(f"tile {tile}: start stamp {t0 - t[0, 23, 6]}")
    prev = t0
    for s in range(24):
        r = t[tile, s]
        if r[0] == 0:
            continue
        f = lambda x: (x - t0) if x else -1
        print(f"  step {s:2d}: mma_saw {f(r[0]):6d} issued {f(r[1]):6d} (+{r[1]-r[0]:4d}) | epi_saw {f(r[2]):6d} (+{(r[2]-r[1]) if r[2] else 0:5d}) ld {f(r[3]):6d} (+{(r[3]-r[2]) if r[3] else 0:4d}) "
              f"written {f(r[4]):6d} (+{(r[4]-r[3]) if r[4] else 0:4d}) arrived {f(r[5]):6d} (+{(r[5]-r[4]) if r[5] else 0:4d})  step total {((r[5] or r[1]) - prev):6d}")
        prev = r[5] or r[1]
