#!/usr/bin/env python
"""Convert weights saved by the reference (`model.save_weights(path)`, Keras 2.2.4 HDF5; BS_brain.py:863-870) into the
`.npz` this engine's `BS.model.load_weights` reads.  h5py is not part of this image, so the converter is
a script for a machine that has it (any Python 3 with `pip install h5py numpy`):

    python scripts/h5_to_npz.py Q-Network_model_weights-Episode-100-....h5  q_network.npz

Order of the output (keys w000, w001, ...) = the order `BS._get_weight_list` / Keras `get_weights()` of THIS engine uses:
layer by layer (GNN stage 1, 2, 3, then Dense 80, 40, 20, CH), node slot by node slot, [W1, W2, W3, bias] for a GNNLayer
(BS_brain.py:26-41) and [kernel, bias] for a Dense.  Keras files are keyed by layer NAME: stage-1 layers are D{k}_GNN
(:121-142), the unnamed stage-2/3 layers and hidden Dense layers are gnn_layer_<i> / dense_<i> in creation order
(slots 1..4 of stage 2, then of stage 3; per slot Dense 80, 40, 20), the output layers D{k}_Decide_Output (:176-200).
A model saved from the target network (second `_create_model` call) simply has higher counters; sorting by suffix works
for both."""
import re
import sys

import numpy as np


def suffix(name):
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def layer_weights(f, name):
    g = f[name]
    names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs["weight_names"]]
    return [np.asarray(g[n]) for n in names]


def main(src, dst, N=4):
    import h5py
    with h5py.File(src, "r") as f:
        root = f["model_weights"] if "model_weights" in f else f
        names = [n.decode() if isinstance(n, bytes) else n for n in root.attrs["layer_names"]]
        have = [n for n in names if len(root[n].attrs["weight_names"])]
        gnn_auto = sorted([n for n in have if re.match(r"gnn_layer_\d+$", n)], key=suffix)
        dense_auto = sorted([n for n in have if re.match(r"dense_\d+$", n)], key=suffix)
        assert len(gnn_auto) == 2 * N and len(dense_auto) == 3 * N, (gnn_auto, dense_auto)
        out = []
        for stage_names in ([f"D{k + 1}_GNN" for k in range(N)], gnn_auto[:N], gnn_auto[N:]):
            for n in stage_names:
                ws = layer_weights(root, n)                 # Keras stores them in add_weight order: W1, W2, W3, bias
                assert len(ws) == 4, (n, [w.shape for w in ws])
                out += ws
        for j in range(3):
            for k in range(N):
                out += layer_weights(root, dense_auto[3 * k + j])
        for k in range(N):
            out += layer_weights(root, f"D{k + 1}_Decide_Output")
    np.savez(dst, **{f"w{i:03d}": w.astype(np.float32) for i, w in enumerate(out)})
    print(f"{dst}: {len(out)} arrays, {sum(w.size for w in out)} parameters")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
