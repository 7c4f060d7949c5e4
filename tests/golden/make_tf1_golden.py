#!/usr/bin/env python
"""Pin the brain oracle against the REAL reference: golden vectors from the unmodified BS_brain.py under Keras 2.2.4 /
TensorFlow 1.14.0 (the versions /root/reference/README.md:9-11 names).

This image has neither package (no cp312 wheel, no network), which is why oracle/v2v_oracle.py says "parity unpinned".
Run this script on any machine that has them (Python 3.6/3.7):

    pip install keras==2.2.4 tensorflow==1.14.0 numpy
    python tests/golden/make_tf1_golden.py --reference /path/to/Globecom2020-ResourceAllocationGNN --out tests/golden

It imports the reference's own `BS` class (BS_brain.py:90-239, nothing re-implemented), injects seeded weights layer by
layer, feeds seeded inputs in the reference's dict format (BS_brain.py:495-504: per-slot arrays + kron(Adj, I_F)) and
writes tests/golden/tf1_n4_*.npz holding inputs, injected weights (engine flat layout), `predict` of both networks,
the `fit` loss (total and per head) and the weights after ONE `train_dnn` step.  tests/test_tf1_golden.py consumes the
files when they exist: the NumPy oracle (CPU) and the CUDA engine (GPU) must both reproduce them to 1e-4 / Adam 1e-6.

Layer <-> slot mapping.  Stage-1 layers are named D{k}_GNN (BS_brain.py:121-142); the stage-2/3 GNNLayers and the hidden
Dense layers are unnamed, so Keras numbers them in creation order (K.get_uid): per model, gnn_layer_* sorted by suffix
= stage 2 slots 1..4 then stage 3 slots 1..4 (:154-164); dense_* sorted by suffix = slots 1..4 x (80, 40, 20) (:176-200);
the output layers are named D{k}_Decide_Output.  Shapes are asserted.
"""
import argparse
import os
import random
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import v2v_oracle as O      # noqa: E402  (NumPy only: dimensions and the flat parameter layout)

N, F, CH, S = 4, 16, 4, 3


def suffix(name):
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def slot_layers(model):
    """[(stage-or-mlp index l, slot k, keras layer)] in the engine's layer order (GNN stages, then the 4 Dense layers)."""
    gnn_named = {k: model.get_layer(f"D{k + 1}_GNN") for k in range(N)}
    gnn_auto = sorted([l for l in model.layers if l.__class__.__name__ == "GNNLayer" and not re.match(r"D\d_GNN$", l.name)],
                      key=lambda l: suffix(l.name))
    dense_auto = sorted([l for l in model.layers if l.__class__.__name__ == "Dense" and not l.name.endswith("_Decide_Output")],
                        key=lambda l: suffix(l.name))
    assert len(gnn_auto) == 2 * N and len(dense_auto) == 3 * N, (len(gnn_auto), len(dense_auto))
    out = []
    for k in range(N):
        out.append((0, k, gnn_named[k]))
        out.append((1, k, gnn_auto[k]))
        out.append((2, k, gnn_auto[N + k]))
        for j in range(3):
            out.append((S + j, k, dense_auto[3 * k + j]))
        out.append((S + 3, k, model.get_layer(f"D{k + 1}_Decide_Output")))
    return out


def inject(model, dims, layers):
    """layers: oracle structure [{'W': [G,K,O], 'b': [G,O]}]; GNN weights are split [W1; W2; W3] (BS_brain.py:26-37)."""
    for l, k, layer in slot_layers(model):
        W, b = layers[l]["W"][k].astype(np.float32), layers[l]["b"][k].astype(np.float32)
        if l < S:
            da = dims.Dn if l == 0 else dims.F + dims.Dn
            ws = [W[:da], W[da:da + dims.De], W[da + dims.De:], b]
        else:
            ws = [W, b]
        shapes = [tuple(w.shape) for w in layer.get_weights()]
        assert shapes == [tuple(w.shape) for w in ws], (layer.name, shapes, [w.shape for w in ws])
        layer.set_weights(ws)


def extract(model, dims, like):
    layers = [{"W": np.zeros_like(l["W"]), "b": np.zeros_like(l["b"])} for l in like]
    for l, k, layer in slot_layers(model):
        ws = layer.get_weights()
        if l < S:
            layers[l]["W"][k] = np.concatenate(ws[:3], 0)
            layers[l]["b"][k] = ws[3]
        else:
            layers[l]["W"][k], layers[l]["b"][k] = ws
    return layers


def generate(reference, out, batches, prefix="tf1", require_pinned_stack=True):
    """Drive the reference's own BS class (imported from `reference`) and write {prefix}_n4_b{B}.npz for every B."""
    sys.path.insert(0, reference)
    import keras                                  # noqa: F401  (fails loudly if the stack is missing)
    import tensorflow as tf
    if require_pinned_stack:
        assert keras.__version__.startswith("2.2") and tf.__version__.startswith("1.14") and "shim" not in keras.__version__, \
            (keras.__version__, tf.__version__)
    from BS_brain import BS                       # the reference's own class, unmodified
    dims = O.BrainDims(N, 3, 1, F, 1, CH, stages=S, per_slot=True)
    for B in batches:
        rng = np.random.default_rng(1001 + B)     # the reference's training seed (RL_Train_main.py:44) + batch
        np.random.seed(1001 + B)                  # Keras' fit shuffles with the global generator (RL_Train_main.py:45-47)
        random.seed(1001 + B)
        brain = BS(N, 3, 1, F, 1, CH)
        online = O.init_params(dims, rng, dtype=np.float32, bias_scale=0.05)
        target = O.init_params(dims, rng, dtype=np.float32, bias_scale=0.05)
        inject(brain.model, dims, online)
        inject(brain.target_model, dims, target)
        node, edge, adj, _ = O.synth_batch(B, N, rng)
        node, edge = node.astype(np.float32), edge.astype(np.float32)
        x = {"Adjacency_Matrix": np.kron(adj, np.eye(F))}
        for k in range(N):
            x[f"D{k + 1}_Node_Input"] = node[:, k].astype(np.float64)
            x[f"D{k + 1}_Edge_Input"] = edge[:, k].astype(np.float64)
            x[f"D{k + 1}_Neighbor_Input"] = np.zeros((B, F))
        q = np.stack(brain.predict(x), 1)
        q_t = np.stack(brain.predict(x, target=True), 1)
        actions = rng.integers(0, CH, (B, N))
        rewards = rng.normal(10.0, 3.0, B)
        y = O.td_targets(q.astype(np.float64), q_t.astype(np.float64), actions, rewards, 0.5).astype(np.float32)   # :668-692
        hist = brain.train_dnn(x, {f"D{k + 1}_Decide_Output": y[:, k] for k in range(N)}, B)
        per_head = np.array([hist.history[f"D{k + 1}_Decide_Output_loss"][0] for k in range(N)])
        after = extract(brain.model, dims, online)
        brain.update_target_model()
        synced = extract(brain.target_model, dims, online)
        assert all(np.array_equal(s["W"], t["W"]) for s, t in zip(synced, after))
        np.savez_compressed(os.path.join(out, f"{prefix}_n4_b{B}.npz"), N=N, S=S, per_slot=1, F=F, CH=CH,
                            node=node, edge=edge, adj=adj.astype(np.float32), params=O.flatten_params(online),
                            target_params=O.flatten_params(target), q=q, q_target=q_t, actions=actions, rewards=rewards,
                            y=y, loss=float(hist.history["loss"][0]), per_head=per_head,
                            params_after_fit=O.flatten_params(after), keras=keras.__version__, tensorflow=tf.__version__)
        print(f"wrote {prefix}_n4_b{B}.npz: loss {hist.history['loss'][0]:.6f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of Coolzyh/Globecom2020-ResourceAllocationGNN")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 64, 256])
    a = ap.parse_args()
    generate(a.reference, a.out, a.batches)


if __name__ == "__main__":
    main()
