"""bf16 configuration of the brain (BASELINE configs[2]): the tcgen05 training kernel (csrc/tc_train.cu) against

  (a) oracle/bf16_emul.py -- the fp64 oracle with the kernel's rounding points made explicit: TIGHT tolerance
      (only accumulation order and rounding-boundary flips differ), and
  (b) oracle/v2v_oracle.py in fp64 -- the arithmetic the reference states: LOOSE tolerance, the price of bf16 operands
      (8 significand bits; north_star's 1e-4 bar is the fp32 configuration's, checked in test_gpu_brain.py).

Tolerances (|x - ref| / max |ref|).  Q vs (a): median <= 1e-6, at most 3 % of the elements beyond 1e-3 and none beyond
2e-2 -- the device accumulates in fp32 (order unspecified) where the emulation accumulates in fp64, so a value that sits
on a bf16 rounding boundary can flip by one bf16 ulp (2^-8 relative; the aggregated rows reach |a| ~ 2e3, ulp 16) and
that flip propagates to ~1 % of the outputs; everything else agrees to ~1e-8.  Q vs (b): 5e-2.  Gradients vs (a) 5e-3,
vs (b) 0.15 with cosine similarity >= 0.99 on well-conditioned targets (a consistent TD-like shift; with pure-noise targets the gradient is a
random-sign sum and bf16-vs-fp64 is ill-conditioned whatever the implementation: measured 0.3 for the emulation itself;
the emulation's own distance to fp64 on these cases is 1-3e-2 for Q and 3-9e-2 for the gradients).
"""
import numpy as np
import pytest
import torch

from oracle import v2v_oracle as O
from oracle import bf16_emul as E

pytestmark = pytest.mark.gpu

Q_EMUL_MAX, Q_F64, G_EMUL, G_F64 = 2e-2, 5e-2, 5e-3, 0.15


def check_q_against_emulation(q, q_emul):
    err = np.abs(np.asarray(q, np.float64) - q_emul) / np.abs(q_emul).max()
    assert np.median(err) <= 1e-6, np.median(err)
    assert (err > 1e-3).mean() <= 0.03, (err > 1e-3).mean()
    assert err.max() <= Q_EMUL_MAX, err.max()


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def setup(v2v, N, S, B, seed):
    rng = np.random.default_rng(seed)
    d = O.BrainDims(N, stages=S, per_slot=False)
    L = O.init_params(d, rng, bias_scale=0.05)
    for l in L:
        l["W"], l["b"] = l["W"].astype(np.float32).astype(np.float64), l["b"].astype(np.float32).astype(np.float64)
    node, edge, adj, _ = O.synth_batch(B, N, rng)
    node, edge = node.astype(np.float32), edge.astype(np.float32)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, dtype="bf16")
    brain.set_flat_params(O.flatten_params(L), 0)
    return rng, d, L, node, edge, adj, brain


@pytest.mark.parametrize("N,S,B", [(20, 3, 64), (20, 2, 200), (20, 3, 1000), (4, 3, 257), (7, 1, 50), (32, 2, 77), (2, 2, 9)])
def test_bf16_forward(v2v, N, S, B):
    rng, d, L, node, edge, adj, brain = setup(v2v, N, S, B, 300 + N + S)
    x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
    q = np.stack(brain.predict(x), 1)
    q_emul = E.brain_forward_backward_bf16(d, L, node.astype(np.float64), edge.astype(np.float64), adj)
    q_f64 = O.brain_forward(d, L, node.astype(np.float64), edge.astype(np.float64), adj)
    check_q_against_emulation(q, q_emul)
    assert rel(q, q_f64) <= Q_F64, rel(q, q_f64)
    # target network and the device entry point
    brain.update_target_model()
    dev = lambda a: torch.from_numpy(a).cuda()
    im, _, _ = v2v.pack_adjacency(dev(adj.astype(np.float32)))
    qd = brain.forward_device(dev(node), dev(edge), in_mask=im, target=True).cpu().numpy()
    assert np.array_equal(qd, q)                                  # same weights, same kernel: bit-identical


@pytest.mark.parametrize("N,S,B", [(20, 3, 64), (20, 3, 1000), (20, 2, 1024), (4, 3, 257), (7, 1, 50), (32, 2, 77)])
def test_bf16_train_step(v2v, N, S, B):
    rng, d, L, node, edge, adj, brain = setup(v2v, N, S, B, 400 + N + S)
    x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
    q = np.stack(brain.predict(x), 1).astype(np.float64)
    y = (q + 0.4 + rng.normal(0, 0.3, q.shape)).astype(np.float32)      # TD-like consistent shift (see the module docstring)
    f64 = lambda a: a.astype(np.float64)
    _, loss_e, ph_e, g_e = E.brain_forward_backward_bf16(d, L, f64(node), f64(edge), adj, f64(y), q_for_loss=q)
    loss_o, ph_o, g_o = O.brain_backward(d, L, f64(node), f64(edge), adj, f64(y), q_for_loss=q)
    p0 = brain.get_flat_params(0).astype(np.float64)
    h = brain.train_dnn(x, {"Decide_Output": y}, B)
    assert abs(h.history["loss"][0] - loss_e) <= 1e-4 * abs(loss_e)
    for k in range(N):
        assert abs(h.history[f"D{k + 1}_Decide_Output_loss"][0] - ph_e[k]) <= 1e-4 * max(ph_e.max(), 1e-9)
    g = brain.get_flat_params(2).astype(np.float64)
    ge, go = O.flatten_params(g_e), O.flatten_params(g_o)
    assert rel(g, ge) <= G_EMUL, rel(g, ge)
    assert rel(g, go) <= G_F64, rel(g, go)
    cos = float(g @ go / (np.linalg.norm(g) * np.linalg.norm(go)))
    assert cos >= 0.99, cos
    # dead parameters (stage-0 neighbour rows: the reference feeds zeros, BS_brain.py:478) have exactly zero gradient
    g_layers = O.unflatten_params(d, g)
    assert np.all(g_layers[0]["W"][0, d.Dn + d.De:] == 0.0)
    # fp32 master weights + Keras-Adam on the device's gradient, exactly
    p1, m1, v1 = O.keras_adam_step(p0, g, np.zeros_like(p0), np.zeros_like(p0), 1)
    assert np.abs(brain.get_flat_params(0) - p1).max() <= 1e-6
    assert brain.iterations == 1
    # deterministic: the same step from the same state gives bit-identical gradients
    brain.set_flat_params(p0.astype(np.float32), 0)
    brain.train_dnn(x, {"Decide_Output": y}, B)
    assert np.array_equal(brain.get_flat_params(2).astype(np.float64), g)


def test_bf16_device_step_equals_host_step(v2v):
    N, S, B = 20, 3, 512
    rng, d, L, node, edge, adj, brain = setup(v2v, N, S, B, 77)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nd, ed = dev(node), dev(edge)
    im, om, binary = v2v.pack_adjacency(dev(adj.astype(np.float32)))
    assert binary
    q = brain.forward_device(nd, ed, in_mask=im)
    y = (q + 0.5).contiguous()
    p0 = brain.get_flat_params(0)
    l_dev = brain.train_step_device(nd, ed, im, om, None, y).cpu().numpy()
    # huber(0.5) = 0.125 per element -> 0.125 per head (mean over B x CH); |q| ~ 1e3 in fp32 leaves ~1e-4 on q + 0.5 - q
    assert np.abs(l_dev - 0.125).max() <= 1e-3 * 0.125
    g_dev, p_dev = brain.get_flat_params(2), brain.get_flat_params(0)
    brain.set_flat_params(p0, 0)
    brain.set_flat_params(np.zeros_like(p0), 3); brain.set_flat_params(np.zeros_like(p0), 4)
    brain._lib.v2v_brain_set_iterations(brain._handle, 0)
    x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
    h = brain.train_dnn(x, {"Decide_Output": y.cpu().numpy()}, B)
    assert np.array_equal(brain.get_flat_params(2), g_dev) and np.array_equal(brain.get_flat_params(0), p_dev)
    assert abs(h.history["loss"][0] - float(l_dev.sum())) <= 1e-6 * float(l_dev.sum())
    for _ in range(5):                                            # further steps stay finite
        assert torch.isfinite(brain.train_step_device(nd, ed, im, om, None, y)).all()
    assert brain.iterations == 6


def test_bf16_rejections(v2v):
    with pytest.raises(ValueError):
        v2v.BS(4, 3, 1, 16, 1, 4, per_slot=True, data_parallel=False, dtype="bf16")          # per-slot weights
    with pytest.raises(ValueError):
        v2v.BS(20, 3, 1, 16, 1, 4, stages=4, per_slot=False, data_parallel=False, dtype="bf16")
    with pytest.raises(ValueError):
        v2v.BS(20, 3, 1, 16, 1, 4, per_slot=False, data_parallel=False, dtype="fp8")
    brain = v2v.BS(4, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=8, data_parallel=False, dtype="bf16", seed=1)
    rng = np.random.default_rng(0)
    node, edge, adj, _ = O.synth_batch(8, 4, rng)
    x = {"Node_Input": node.astype(np.float32), "Edge_Input": edge.astype(np.float32), "Adjacency_Matrix": 0.5 * adj}
    with pytest.raises(ValueError):
        brain.predict(x)                                          # weighted adjacency: fp32 configuration only
