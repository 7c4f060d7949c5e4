"""bf16 tensor-core brain against oracle/bf16_emul.py: where do the deviations sit? (scratch; a checker like tests/)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from oracle import v2v_oracle as O, bf16_emul as E
N, S, B = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (20, 3, 64)))
rng = np.random.default_rng(300 + N + S)
d = O.BrainDims(N, stages=S, per_slot=False)
L = O.init_params(d, rng, bias_scale=0.05)
for l in L:
    l["W"], l["b"] = l["W"].astype(np.float32).astype(np.float64), l["b"].astype(np.float32).astype(np.float64)
node, edge, adj, _ = O.synth_batch(B, N, rng)
node, edge = node.astype(np.float32), edge.astype(np.float32)
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, dtype="bf16")
brain.set_flat_params(O.flatten_params(L), 0)
x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
q = np.stack(brain.predict(x), 1).astype(np.float64)
qe = E.brain_forward_backward_bf16(d, L, node.astype(np.float64), edge.astype(np.float64), adj)
err = np.abs(q - qe) / np.abs(qe).max()
print("max rel", err.max(), "mean rel", err.mean(), "median", np.median(err))
print("per graph max (first 16):", np.round(err.max(axis=(1, 2))[:16], 5))
print("per node max:", np.round(err.max(axis=(0, 2)), 5))
print("per channel max:", np.round(err.max(axis=(0, 1)), 5))
print("fraction of elements > 1e-3:", (err > 1e-3).mean(), " > 1e-4:", (err > 1e-4).mean())
b, n, c = np.unravel_index(err.argmax(), err.shape)
print("worst at", b, n, c, q[b, n, c], qe[b, n, c])
# train step
y = (q + 0.4 + rng.normal(0, 0.3, q.shape)).astype(np.float32)
f64 = lambda a: a.astype(np.float64)
_, loss_e, ph_e, g_e = E.brain_forward_backward_bf16(d, L, f64(node), f64(edge), adj, f64(y), q_for_loss=q)
h = brain.train_dnn(x, {"Decide_Output": y}, B)
print("loss", h.history["loss"][0], loss_e)
g = brain.get_flat_params(2).astype(np.float64)
gl = O.unflatten_params(d, g)
for li, (a, b_) in enumerate(zip(gl, g_e)):
    for k in ("W", "b"):
        ref = b_[k]; got = a[k]
        print(f"layer {li} {k}: max|ref| {np.abs(ref).max():.3e}  max|diff| {np.abs(got - ref).max():.3e}  rel {np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30):.3e}")
