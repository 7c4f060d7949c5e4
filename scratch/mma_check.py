"""Tensor-core (mma.sync 3xTF32) modes of the fused fp32 kernel against the FP32-pipe mode and the layered kernels:
forward / loss / gradient deviations and step times (V2V_FUSED_MMA modes 0, 1)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
lib = v2v.load_library()
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
def run(N, S, B, seed=0):
    rng = np.random.default_rng(seed)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, seed=9)
    node, edge, adj = synth_numpy(B, N, rng)
    nd, ed, ad = (torch.from_numpy(t.astype(np.float32)).cuda() for t in (node, edge, adj))
    im, om, _ = v2v.pack_adjacency(ad)
    p0 = brain.get_flat_params(0); p0 += rng.normal(0, 0.02, p0.shape).astype(np.float32)
    res = {}
    y = None
    for mode in ("layered", 0, 1):
        if mode == "layered": brain.set_fused(0)
        else:
            brain.set_fused(1); assert lib.v2v_fused_set_mma(mode) == 0
        brain.set_flat_params(p0, 0)
        for w in (3, 4): brain.set_flat_params(np.zeros_like(p0), w)
        v2v._lib.check(lib.v2v_brain_set_iterations(brain._handle, 0))
        q = brain.forward_device(nd, ed, in_mask=im).cpu().numpy()
        if y is None: y = torch.from_numpy((q + rng.normal(0, 1.0, q.shape)).astype(np.float32)).cuda()
        hl = brain.train_step_device(nd, ed, im, om, None, y).cpu().numpy()
        res[mode] = (q, hl, brain.get_flat_params(2), brain.get_flat_params(0))
    ql, hl_, gl, pl = res["layered"]
    for mode in (0, 1):
        q, hl, g, p = res[mode]
        print(f"N={N} S={S} B={B} mode {mode}: q {rel(q, ql):.2e} loss {rel(hl, hl_):.2e} grad {rel(g, gl):.2e} (vs mode0 {rel(g, res[0][2]):.2e}) "
              f"|q|max {np.abs(ql).max():.0f} dp max {np.abs(p - pl).max():.2e} q90 {np.quantile(np.abs(p - pl), 0.9):.2e} nan {int(np.isnan(g).sum())}", flush=True)
    return brain, (nd, ed, im, om, y)
for cfg in ((20, 2, 1024), (20, 3, 333), (4, 3, 64), (7, 2, 100), (32, 2, 200), (9, 1, 50)):
    run(*cfg)
# multi-tile per CTA + timing
for B in (1024, 8192):
    brain, (nd, ed, im, om, y) = run(20, 2, B, seed=1)
    brain.set_fused(1)
    for mode in (0, 1):
        lib.v2v_fused_set_mma(mode)
        for _ in range(5): brain.train_step_device(nd, ed, im, om, None, y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): brain.train_step_device(nd, ed, im, om, None, y)
        e1.record(); torch.cuda.synchronize()
        t_tr = 10 * e0.elapsed_time(e1)
        q = brain.forward_device(nd, ed, in_mask=im)
        e0.record()
        for _ in range(100): brain.forward_device(nd, ed, in_mask=im, out=q)
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} mode {mode}: train step {t_tr:.1f} us, forward {10 * e0.elapsed_time(e1):.1f} us", flush=True)
