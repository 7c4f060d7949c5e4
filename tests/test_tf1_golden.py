"""Real-reference pins: vectors written by tests/golden/make_tf1_golden.py from the UNMODIFIED BS_brain.py under Keras 2.2.4 /
TensorFlow 1.14.0.  That stack cannot be installed in this image, so the files are produced elsewhere and
dropped into tests/golden/tf1_*.npz; until they exist these tests skip and the brain oracle stays "parity unpinned".
When they exist: the NumPy oracle (CPU) and the CUDA engine (GPU) must both reproduce the reference's predict, loss and
post-fit weights -- 1e-4 relative (north_star), Adam step 2e-6 absolute (fp32 TF arithmetic on ~1e-3 updates)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import v2v_oracle as O

FILES = sorted(glob.glob(os.path.join(GOLDEN, "tf1_*.npz")))
needs_files = pytest.mark.skipif(not FILES, reason="no tests/golden/tf1_*.npz (run tests/golden/make_tf1_golden.py where Keras 2.2.4 / "
                                                   "TF 1.14 exist)")


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@needs_files
@pytest.mark.parametrize("path", FILES or [None])
def test_oracle_reproduces_tf1(path):
    z = np.load(path)
    d = O.BrainDims(int(z["N"]), stages=int(z["S"]), per_slot=bool(z["per_slot"]))
    f64 = lambda k: np.asarray(z[k], np.float64)
    L, Lt = O.unflatten_params(d, f64("params")), O.unflatten_params(d, f64("target_params"))
    assert rel(O.brain_forward(d, L, f64("node"), f64("edge"), f64("adj")), z["q"]) <= 1e-4
    assert rel(O.brain_forward(d, Lt, f64("node"), f64("edge"), f64("adj")), z["q_target"]) <= 1e-4
    loss, per_head, g = O.brain_backward(d, L, f64("node"), f64("edge"), f64("adj"), f64("y"), q_for_loss=f64("q"))
    assert abs(loss - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    assert rel(per_head, z["per_head"]) <= 1e-4
    p1, _, _ = O.keras_adam_step(f64("params"), O.flatten_params(g), 0.0, 0.0, 1)
    assert np.abs(p1 - z["params_after_fit"]).max() <= 2e-6


@needs_files
@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES or [None])
def test_engine_reproduces_tf1(v2v, path):
    z = np.load(path)
    N, B = int(z["N"]), z["node"].shape[0]
    brain = v2v.BS(N, 3, 1, int(z["F"]), 1, int(z["CH"]), stages=int(z["S"]), per_slot=bool(z["per_slot"]), max_batch=B,
                   data_parallel=False)
    brain.set_flat_params(z["params"], 0)
    brain.set_flat_params(z["target_params"], 1)
    x = {"Adjacency_Matrix": np.kron(z["adj"].astype(np.float64), np.eye(int(z["F"])))}
    for k in range(N):
        x[f"D{k + 1}_Node_Input"] = z["node"][:, k].astype(np.float64)
        x[f"D{k + 1}_Edge_Input"] = z["edge"][:, k].astype(np.float64)
        x[f"D{k + 1}_Neighbor_Input"] = np.zeros((B, int(z["F"])))
    assert rel(np.stack(brain.predict(x), 1), z["q"]) <= 1e-4
    assert rel(np.stack(brain.predict(x, target=True), 1), z["q_target"]) <= 1e-4
    h = brain.train_dnn(x, {f"D{k + 1}_Decide_Output": z["y"][:, k] for k in range(N)}, B)
    assert abs(h.history["loss"][0] - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    for k in range(N):
        assert abs(h.history[f"D{k + 1}_Decide_Output_loss"][0] - z["per_head"][k]) <= 1e-4 * z["per_head"].max()
    assert np.abs(brain.get_flat_params(0) - z["params_after_fit"]).max() <= 2e-6
