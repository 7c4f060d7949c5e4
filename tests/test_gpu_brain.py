"""GPU parity of the brain (BS) through the reference-facing Python surface and the C-ABI host
entry points, against the committed golden vectors and the live fp64 oracle.

Tolerance: 1e-4 relative (north_star, fp32) -- measured as max|x - ref| <= 1e-4 * max|ref|.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases
from oracle import v2v_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def ref_dict(node, edge, adj, F=16, kron=True, neighbor=True):
    """The reference's input dict (BS_brain.py:495-504): per-slot fp64 arrays + Kronecker A."""
    B, N = node.shape[:2]
    d = {}
    for k in range(N):
        d[f"D{k + 1}_Node_Input"] = node[:, k].astype(np.float64)
        d[f"D{k + 1}_Edge_Input"] = edge[:, k].astype(np.float64)
        if neighbor:
            d[f"D{k + 1}_Neighbor_Input"] = np.zeros((B, F))
    d["Adjacency_Matrix"] = np.kron(adj, np.eye(F)) if kron else adj
    return d


def check_grads(g_gpu, z, q_gpu, neigh=None):
    """Backward parity.  The Huber residual q - y cancels catastrophically when |q| >> |q - y|
    (untrained nets at N=20 output |q| ~ 1e3 against residuals ~ 1): ANY fp32 forward, the
    reference's TF1 graph included, then carries ~eps*|q| absolute error into dq.  So (1) the
    backward kernels are checked at 1e-4 against the fp64 oracle fed with the device's own
    fp32 output inside the residual, and (2) the whole chain against the pure-fp64 gradient with
    that measured sensitivity added to the tolerance."""
    d = O.BrainDims(int(z["N"]), stages=int(z["S"]), per_slot=bool(z["per_slot"]))
    L = O.unflatten_params(d, np.asarray(z["params"], np.float64))
    args = [np.asarray(z[k], np.float64) for k in ("node", "edge", "adj", "y")]
    _, _, g_cond = O.brain_backward(d, L, *args, neigh=neigh, q_for_loss=q_gpu)
    g_cond = O.flatten_params(g_cond)
    assert rel_err(g_gpu, g_cond) <= RTOL
    sens = rel_err(g_cond, z["grads"])
    assert rel_err(g_gpu, z["grads"]) <= RTOL + 2 * sens


def make_brain(v2v, z, **kw):
    N, S, per_slot = int(z["N"]), int(z["S"]), bool(z["per_slot"])
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=per_slot, max_batch=64, data_parallel=False, **kw)
    brain.set_flat_params(z["params"], 0)
    brain.set_flat_params(z["target_params"], 1)
    return brain


@pytest.mark.parametrize("name", golden_cases())
def test_golden_predict_train(v2v, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    brain = make_brain(v2v, z)
    N = int(z["N"])
    x = ref_dict(z["node"], z["edge"], z["adj"], kron=(N <= 8))
    # --- predict: online and target nets
    p = brain.predict(x)
    assert isinstance(p, list) and len(p) == N and p[0].shape == (z["node"].shape[0], 4)
    q_gpu = np.stack(p, 1).astype(np.float64)
    assert rel_err(q_gpu, z["q"]) <= RTOL
    p_t = brain.predict(x, target=True)
    assert rel_err(np.stack(p_t, 1), z["q_target"]) <= RTOL
    # --- the agent mutates the returned arrays in place (BS_brain.py:684-690): they must be writable and independent
    p[0][0, 0] = 123.0
    assert p[1][0, 0] != 123.0 and brain.predict(x)[0][0, 0] != 123.0
    # --- one fit step: loss, per-head losses and gradients
    y = {f"D{k + 1}_Decide_Output": z["y"][:, k] for k in range(N)}
    B = z["node"].shape[0]
    h = brain.train_dnn(x, y, B)
    assert abs(h.history["loss"][0] - float(z["loss"])) <= RTOL * abs(float(z["loss"]))
    for k in range(N):
        assert abs(h.history[f"D{k + 1}_Decide_Output_loss"][0] - z["per_head"][k]) <= RTOL * max(z["per_head"].max(), 1e-9)
    check_grads(brain.get_flat_params(2), z, q_gpu)
    assert brain.iterations == 1
    # --- Keras-Adam.  The rule is m/(sqrt(v)+1e-7): for the handful of weights whose gradient is
    # ~1e-7 the update is as ill-conditioned as the fp32 gradient itself, so the optimiser kernel is
    # checked exactly (fp64 rule applied to the device's own gradients), the loss of every step
    # against the oracle evaluated at the device's parameters, and the committed fp64 trajectory in
    # bulk (Adam normalises each weight's step to ~lr whatever its gradient's size, so weights with
    # |g| << max|g| amplify the fp32 rounding of g: mean drift <= 1e-4, none beyond three full steps).
    d = O.BrainDims(N, stages=int(z["S"]), per_slot=bool(z["per_slot"]))
    arrs = [np.asarray(z[k], np.float64) for k in ("node", "edge", "adj", "y")]
    p_ref = z["params"].astype(np.float64)
    m_ref, v_ref = np.zeros_like(p_ref), np.zeros_like(p_ref)
    p_ref, m_ref, v_ref = O.keras_adam_step(p_ref, brain.get_flat_params(2).astype(np.float64), m_ref, v_ref, 1)
    assert np.abs(brain.get_flat_params(0) - p_ref).max() <= 1e-6
    for t in (2, 3):
        p_now = brain.get_flat_params(0).astype(np.float64)
        loss_ref, _, _ = O.brain_backward(d, O.unflatten_params(d, p_now), *arrs)
        loss_t = brain.train_dnn(x, y, B).history["loss"][0]
        assert abs(loss_t - loss_ref) <= RTOL * abs(loss_ref)
        p_ref, m_ref, v_ref = O.keras_adam_step(p_now, brain.get_flat_params(2).astype(np.float64), m_ref, v_ref, t)
        assert np.abs(brain.get_flat_params(0) - p_ref).max() <= 1e-6
        assert np.abs(brain.get_flat_params(3) - m_ref).max() <= 1e-6 * max(1.0, np.abs(m_ref).max())
    assert brain.iterations == 3
    diff = np.abs(brain.get_flat_params(0) - z["params_after_adam"])
    assert diff.mean() <= 1e-4 and diff.max() <= 3.2e-3
    # target net untouched by training, then synchronised
    assert np.array_equal(brain.get_flat_params(1), z["target_params"])
    brain.update_target_model()
    assert np.array_equal(brain.get_flat_params(1), brain.get_flat_params(0))


@pytest.mark.parametrize("N,S,per_slot,B,kind", [
    (20, 2, False, 1024, None), (20, 3, False, 257, 2), (20, 3, True, 64, None), (4, 3, True, 512, None),
    (4, 3, True, 1, None), (9, 1, False, 33, None), (32, 3, False, 40, None), (40, 2, False, 20, None),
])
def test_forward_backward_vs_live_oracle(v2v, N, S, per_slot, B, kind):
    rng = np.random.default_rng(N * 1000 + S * 10 + B)
    d = O.BrainDims(N, stages=S, per_slot=per_slot)
    L = O.init_params(d, rng, bias_scale=0.05)
    for l in L:
        l["W"], l["b"] = l["W"].astype(np.float32).astype(np.float64), l["b"].astype(np.float32).astype(np.float64)
    node, edge, adj, _ = O.synth_batch(B, N, rng, sparse_in_degree=kind)
    node, edge = node.astype(np.float32), edge.astype(np.float32)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=per_slot, max_batch=16, data_parallel=False)
    brain.set_flat_params(O.flatten_params(L), 0)
    x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
    q = np.stack(brain.predict(x), 1)
    qr = O.brain_forward(d, L, node.astype(np.float64), edge.astype(np.float64), adj)
    assert rel_err(q, qr) <= RTOL
    y = (qr + rng.normal(0, 1.5, qr.shape)).astype(np.float32)
    loss, per_head, g = O.brain_backward(d, L, node.astype(np.float64), edge.astype(np.float64), adj, y.astype(np.float64))
    h = brain.train_dnn(x, {"Decide_Output": y}, B)
    assert abs(h.history["loss"][0] - loss) <= RTOL * abs(loss)
    z = {"N": N, "S": S, "per_slot": per_slot, "params": O.flatten_params(L), "node": node, "edge": edge, "adj": adj,
         "y": y, "grads": O.flatten_params(g)}
    check_grads(brain.get_flat_params(2), z, q.astype(np.float64))


@pytest.mark.parametrize("N,S,B,per_slot", [(20, 2, 1024, False), (20, 2, 7, False), (20, 2, 1, False), (20, 3, 333, False),
                                            (4, 3, 256, False), (8, 1, 100, False), (31, 2, 50, False), (20, 2, 2500, False),
                                            # the reference's own model: one weight set per node slot (BS_brain.py:121-200)
                                            (4, 3, 256, True), (4, 3, 1, True), (4, 3, 5, True), (4, 3, 3000, True),
                                            (7, 2, 333, True), (8, 3, 64, True), (2, 1, 9, True)])
def test_fused_kernel_matches_layered_kernels(v2v, N, S, B, per_slot):
    try:
        for mma in ((0, 1) if not per_slot else (1,)):
            _fused_vs_layered(v2v, N, S, B, per_slot, mma)
    finally:
        assert v2v.load_library().v2v_fused_set_mma(1) == 0     # the default


@pytest.mark.parametrize("N,S,B", [(20, 2, 1024), (20, 3, 2100), (4, 3, 64), (9, 1, 50), (32, 2, 200), (20, 2, 8192)])
def test_tensor_core_backward_matches_fp32_pipe(v2v, N, S, B):
    """Backward contractions on the tensor cores (mma.sync TF32, three passes per product) against the same kernel with
    every contraction on the FP32 pipe: identical forward and loss (bit for bit), gradients to 1e-5 of their scale --
    also when a CTA owns several tiles (its later tiles add to the partial row its earlier tiles wrote)."""
    lib = v2v.load_library()
    rng = np.random.default_rng(N + S + B)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, seed=5)
    node, edge, adj, _ = O.synth_batch(B, N, rng)
    nd, ed, ad = (torch.from_numpy(t.astype(np.float32)).cuda() for t in (node, edge, adj))
    im, om, _ = v2v.pack_adjacency(ad)
    p0 = brain.get_flat_params(0) + rng.normal(0, 0.02, brain.get_flat_params(0).shape).astype(np.float32)
    q = None
    res = {}
    try:
        assert lib.v2v_fused_set_mma(2) != 0 and lib.v2v_fused_set_mma(-1) != 0      # only 0 and 1 exist
        for mma in (0, 1):
            assert lib.v2v_fused_set_mma(mma) == 0
            brain.set_flat_params(p0, 0)
            for w in (3, 4):
                brain.set_flat_params(np.zeros_like(p0), w)
            v2v._lib.check(lib.v2v_brain_set_iterations(brain._handle, 0))
            if q is None:
                brain.set_tensor_core(0)
                q = brain.forward_device(nd, ed, in_mask=im)
                y = q + torch.from_numpy(rng.normal(0, 1.0, tuple(q.shape)).astype(np.float32)).cuda()
            hl = brain.train_step_device(nd, ed, im, om, None, y).cpu().numpy()
            res[mma] = (hl, brain.get_flat_params(2))
    finally:
        assert lib.v2v_fused_set_mma(1) == 0
    assert np.array_equal(res[0][0], res[1][0])                  # the forward and the Huber head do not change
    g0, g1 = res[0][1], res[1][1]
    assert np.isfinite(g1).all()
    assert rel_err(g1, g0) <= 1e-5
    assert np.array_equal(g1 == 0, g0 == 0)                      # dead parameters (stage-1 neighbour rows) stay exactly zero


def _fused_vs_layered(v2v, N, S, B, per_slot, mma):
    """The one-launch fused network (shared weights, and per-slot weights up to N = 8) against the layer-by-layer
    kernels: forward to 1e-5, gradients, one Adam step."""
    # mma: which pipe runs the backward contractions of the shared-weight kernel (0 FP32 pipe, 1 tensor cores, mma.sync
    # TF32 in three passes per product); the forward is plain fp32 in both, so q and the loss must not move at all
    assert v2v.load_library().v2v_fused_set_mma(mma) == 0 and v2v.load_library().v2v_fused_get_mma() == mma
    rng = np.random.default_rng(N * 100 + S * 10 + B)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=per_slot, max_batch=B, data_parallel=False, seed=9)
    info = brain.fused_info(B)
    assert info["capable"] == 1 and info["smem_bytes"] <= 227 * 1024
    node, edge, adj, _ = O.synth_batch(B, N, rng)
    nd, ed, ad = (torch.from_numpy(t.astype(np.float32)).cuda() for t in (node, edge, adj))
    im, om, _ = v2v.pack_adjacency(ad)
    p0 = brain.get_flat_params(0)
    p0 += rng.normal(0, 0.02, p0.shape).astype(np.float32)       # non-zero biases
    brain.set_flat_params(p0, 0)
    res = {}
    for mode in (1, 0):               # fused one-launch kernel, layer-by-layer kernels
        brain.set_fused(mode)
        brain.set_flat_params(p0, 0)
        for w in (3, 4):
            brain.set_flat_params(np.zeros_like(p0), w)
        v2v._lib.check(brain._lib.v2v_brain_set_iterations(brain._handle, 0))
        q = brain.forward_device(nd, ed, in_mask=im).cpu().numpy()
        y = torch.from_numpy((q + rng.normal(0, 1.0, q.shape)).astype(np.float32)).cuda() if mode == 1 else res[1][4]
        hl = brain.train_step_device(nd, ed, im, om, None, y).cpu().numpy()
        res[mode] = (q, hl, brain.get_flat_params(2), brain.get_flat_params(0), y)
    ql, hl_, gl, pl, _ = res[0]
    for mode in (1,):
        qf, hf, gf, pf, _ = res[mode]
        assert rel_err(qf, ql) <= 1e-5, mode
        assert rel_err(hf, hl_) <= 1e-5, mode
        # gradients inherit the conditioning of q - y (|q| ~ 1e2..1e3 against residuals ~ 1, see check_grads)
        assert rel_err(gf, gl) <= 5e-5 * max(1.0, np.abs(ql).max() / 100.0), mode
        assert np.abs(pf - pl).max() <= 3e-4, mode      # one Adam step of ~1e-3; tiny-|g| weights amplify rounding
        assert np.quantile(np.abs(pf - pl), 0.9) <= 2e-6, mode
    qf = res[1][0]
    # and against the fp64 oracle
    d = O.BrainDims(N, stages=S, per_slot=per_slot)
    L = O.unflatten_params(d, p0.astype(np.float64))
    qr = O.brain_forward(d, L, node.astype(np.float32).astype(np.float64), edge.astype(np.float32).astype(np.float64), adj)
    assert rel_err(qf, qr) <= RTOL


def test_device_resident_path_matches_host_path(v2v):
    rng = np.random.default_rng(77)
    N, B = 20, 300
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=512, data_parallel=False, seed=3)
    node, edge, adj, _ = O.synth_batch(B, N, rng)
    node, edge, adjf = node.astype(np.float32), edge.astype(np.float32), adj.astype(np.float32)
    q_host = np.stack(brain.predict({"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adjf}), 1)
    nd, ed, ad = (torch.from_numpy(t).cuda() for t in (node, edge, adjf))
    im, om, binary = v2v.pack_adjacency(ad)
    assert binary
    q_dev = brain.forward_device(nd, ed, in_mask=im).cpu().numpy()
    assert np.array_equal(q_dev, q_host)
    # weighted-adjacency kernels give the same answer on a 0/1 matrix
    q_w = brain.forward_device(nd, ed, adj=ad).cpu().numpy()
    assert rel_err(q_w, q_host) <= 1e-6
    y = torch.from_numpy((q_host + rng.normal(0, 1, q_host.shape)).astype(np.float32)).cuda()
    p0 = brain.get_flat_params(0)
    l1 = brain.train_step_device(nd, ed, im, om, None, y).cpu().numpy()
    g1 = brain.get_flat_params(2)
    brain.set_flat_params(p0, 0)
    l2 = brain.train_step_device(nd, ed, None, None, ad, y).cpu().numpy()
    assert rel_err(l2, l1) <= 1e-5 and rel_err(brain.get_flat_params(2), g1) <= 1e-4


def test_weighted_adjacency_and_neighbor_input(v2v):
    """Non-0/1 adjacency and a non-zero D{k}_Neighbor_Input: inputs the reference's graph accepts
    (BS_brain.py:119, :144) even though its agent never produces them."""
    rng = np.random.default_rng(5)
    N, B = 4, 19
    d = O.BrainDims(N, stages=3, per_slot=True)
    L = O.init_params(d, rng, bias_scale=0.05)
    for l in L:
        l["W"], l["b"] = l["W"].astype(np.float32).astype(np.float64), l["b"].astype(np.float32).astype(np.float64)
    node, edge, _, _ = O.synth_batch(B, N, rng)
    node, edge = node.astype(np.float32).astype(np.float64), edge.astype(np.float32).astype(np.float64)
    adj = rng.normal(0.5, 0.5, (B, N, N)).astype(np.float32).astype(np.float64)
    neigh = rng.normal(0, 1, (B, N, 16)).astype(np.float32).astype(np.float64)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, max_batch=32, data_parallel=False)
    brain.set_flat_params(O.flatten_params(L), 0)
    x = ref_dict(node, edge, adj)
    for k in range(N):
        x[f"D{k + 1}_Neighbor_Input"] = neigh[:, k]
    A = np.stack([O.kron_adjacency(a, 16) for a in adj])
    qr = np.stack(O.brain_forward_literal(d, L, node, edge, A, neigh=neigh), 1)
    assert rel_err(np.stack(brain.predict(x), 1), qr) <= RTOL
    y = qr + rng.normal(0, 1.0, qr.shape)
    loss, _, g = O.brain_backward(d, L, node, edge, adj, y, neigh=neigh)
    q_gpu = np.stack(brain.predict(x), 1).astype(np.float64)
    h = brain.train_dnn(x, [y[:, k] for k in range(N)], B)
    assert abs(h.history["loss"][0] - loss) <= RTOL * abs(loss)
    z = {"N": N, "S": 3, "per_slot": True, "params": O.flatten_params(L), "node": node, "edge": edge, "adj": adj,
         "y": y, "grads": O.flatten_params(g)}
    check_grads(brain.get_flat_params(2), z, q_gpu, neigh=neigh)


def test_bs_surface_matches_reference(v2v, tmp_path):
    brain = v2v.BS(4, 3, 1, 16, 1, 4, data_parallel=False, seed=1)
    # attributes the Agent reads (BS_brain.py:299, :414-417)
    assert (brain.num_One_Node_Input, brain.num_One_Edge_Input, brain.num_One_D2D_Input, brain.num_D2D_Input,
            brain.num_Feedback) == (9, 4, 13, 68, 16)
    assert brain.model.count_params() == 37824
    ws = brain.model.get_weights()
    assert len(ws) == 4 * (3 * 4 + 4 * 2)                       # 80 weight tensors
    assert [w.shape for w in ws[:4]] == [(9, 16), (4, 16), (16, 16), (16,)]
    assert all(np.all(w == 0) for w in ws if w.ndim == 1)       # Keras zero biases
    lim = np.sqrt(6.0 / (9 + 16))
    assert np.abs(ws[0]).max() <= lim and np.abs(ws[0]).max() > 0.5 * lim      # glorot_uniform
    # online and target nets are initialised independently (:105-106) until synchronised
    assert not np.array_equal(ws[0], brain.target_model.get_weights()[0])
    # save / load round trip through the target model
    path = str(tmp_path / "w")
    brain.model.save_weights(path)
    brain.target_model.load_weights(path)
    assert all(np.array_equal(a, b) for a, b in zip(brain.model.get_weights(), brain.target_model.get_weights()))
    with pytest.raises(ValueError):
        brain.model.set_weights(ws[:-1])
    # Keras-style input validation
    rng = np.random.default_rng(0)
    node, edge, adj, _ = O.synth_batch(3, 4, rng)
    x = ref_dict(node, edge, adj)
    bad = dict(x); del bad["D3_Edge_Input"]
    with pytest.raises(ValueError):
        brain.predict(bad)
    bad = dict(x); bad["D2_Node_Input"] = np.zeros((3, 8))
    with pytest.raises(ValueError):
        brain.predict(bad)
    bad = dict(x); bad["Adjacency_Matrix"] = np.zeros((3, 60, 60))
    with pytest.raises(ValueError):
        brain.predict(bad)
    with pytest.raises(ValueError):
        brain.train_dnn(x, {"D1_Decide_Output": np.zeros((3, 4))}, 3)
    # predict_one_step with B=1 (acting path, :336) and workspace growth beyond max_batch
    one = ref_dict(node[:1], edge[:1], adj[:1])
    assert brain.predict_one_step(one)[0].shape == (1, 4)
    big_n, big_e, big_a, _ = O.synth_batch(3000, 4, rng)
    q_big = brain.predict(ref_dict(big_n, big_e, big_a, kron=False))
    assert q_big[0].shape == (3000, 4)
    assert np.allclose(q_big[2][:1], brain.predict(ref_dict(big_n[:1], big_e[:1], big_a[:1]))[2], rtol=1e-5, atol=1e-6)


def test_fit_minibatches_like_keras(v2v):
    """rows > batch_size: ceil(rows/batch) optimiser steps in one epoch, loss = row-weighted mean."""
    rng = np.random.default_rng(3)
    brain = v2v.BS(4, 3, 1, 16, 1, 4, data_parallel=False, seed=2)
    node, edge, adj, _ = O.synth_batch(10, 4, rng)
    x = ref_dict(node, edge, adj)
    y = [rng.normal(size=(10, 4)) for _ in range(4)]
    h = brain.model.fit(x, y, batch_size=4, epochs=2, shuffle=False)
    assert brain.iterations == 6 and len(h.history["loss"]) == 2
    assert h.history["loss"][1] < h.history["loss"][0] * 1.5


def test_training_trajectory_follows_oracle(v2v):
    """40 consecutive train_dnn steps on fixed data: no "the loss went down" assertion -- every step's loss must equal the
    fp64 oracle's loss evaluated at the device's own parameters before that step, and every update the exact Keras-Adam
    rule applied to the device's own gradient (so the check cannot drift with the trajectory)."""
    rng = np.random.default_rng(4)
    N = 4
    brain = v2v.BS(N, 3, 1, 16, 1, 4, data_parallel=False, seed=5)
    d = O.BrainDims(N, stages=3, per_slot=True)
    node, edge, adj, _ = O.synth_batch(256, N, rng)
    x = ref_dict(node, edge, adj, kron=False)
    y = [rng.normal(1.0, 0.5, (256, 4)) for _ in range(N)]
    y_ref = np.stack(y, 1).astype(np.float32).astype(np.float64)
    node_r, edge_r = (a.astype(np.float32).astype(np.float64) for a in (node, edge))
    m = v = np.zeros(brain.get_flat_params(0).shape, np.float64)
    for t in range(1, 41):
        p_now = brain.get_flat_params(0).astype(np.float64)
        loss_ref, _, _ = O.brain_backward(d, O.unflatten_params(d, p_now), node_r, edge_r, adj, y_ref)
        loss_t = brain.train_dnn(x, y, 256).history["loss"][0]
        assert np.isfinite(loss_t) and abs(loss_t - loss_ref) <= RTOL * abs(loss_ref), (t, loss_t, loss_ref)
        p_ref, m, v = O.keras_adam_step(p_now, brain.get_flat_params(2).astype(np.float64), m, v, t)
        assert np.abs(brain.get_flat_params(0) - p_ref).max() <= 2e-6, t
    assert brain.iterations == 40
