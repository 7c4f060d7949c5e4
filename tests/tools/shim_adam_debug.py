"""Where do the engine's post-fit weights leave the refshim recording, and how large is the gradient there?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import importlib
v2v = importlib.import_module("globecom2020-resourceallocationgnn_b200")
from oracle import v2v_oracle as O
for name in ("refshim_n4_b1", "refshim_n4_b64"):
    z = np.load(f"tests/golden/{name}.npz")
    N, B, F = int(z["N"]), z["node"].shape[0], int(z["F"])
    brain = v2v.BS(N, 3, 1, F, 1, int(z["CH"]), stages=int(z["S"]), per_slot=True, max_batch=B, data_parallel=False)
    brain.set_flat_params(z["params"], 0)
    x = {"Node_Input": z["node"], "Edge_Input": z["edge"], "Adjacency_Matrix": z["adj"]}
    q = np.stack(brain.predict(x), 1)
    h = brain.train_dnn(x, {f"D{k + 1}_Decide_Output": z["y"][:, k] for k in range(N)}, B)
    p1 = brain.get_flat_params(0); g = brain.get_flat_params(2)
    d = O.BrainDims(N, stages=int(z["S"]), per_slot=True)
    f64 = lambda k: np.asarray(z[k], np.float64)
    L = O.unflatten_params(d, f64("params"))
    _, _, go = O.brain_backward(d, L, f64("node"), f64("edge"), f64("adj"), f64("y"), q_for_loss=q.astype(np.float64))
    go = O.flatten_params(go)
    _, _, gq = O.brain_backward(d, L, f64("node"), f64("edge"), f64("adj"), f64("y"), q_for_loss=f64("q"))
    gq = O.flatten_params(gq)
    dev = np.abs(p1 - z["params_after_fit"])
    pa, _, _ = O.keras_adam_step(f64("params"), g.astype(np.float64), 0.0, 0.0, 1)
    print(name, "max |p1 - rec|", dev.max(), " |p1 - adam(p0, g_dev)|", np.abs(pa - p1).max(), " max|g|", np.abs(go).max(),
          " max|g_dev - g_oracle(q_dev)|", np.abs(g - go).max(), " max|g_dev - g_oracle(q_rec)|", np.abs(g - gq).max(),
          " q rel", np.abs(q - z["q"]).max() / np.abs(z["q"]).max())
    for i in np.argsort(-dev)[:8]:
        print(f"   i={i} dp={dev[i]:.2e} g_dev={g[i]:.4e} g_or(q_dev)={go[i]:.4e} g_or(q_rec)={gq[i]:.4e} upd_rec={z['params'][i]-z['params_after_fit'][i]:.3e}")
    print("   elements beyond 2e-6:", int((dev > 2e-6).sum()), "of", dev.size, "; all of them have |g| <", np.abs(gq[dev > 2e-6]).max() if (dev > 2e-6).any() else 0)
