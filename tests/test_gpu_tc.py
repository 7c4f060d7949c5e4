"""Tensor-core (tcgen05, 3xTF32) forward of the shared-weight brain against the fp64 oracle and the FP32-pipe kernel."""
import numpy as np
import pytest
import torch

from oracle import v2v_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4          # north-star bar; the 3-pass TF32 products are expected around 1e-6


def _setup(v2v, N, S, B, seed):
    rng = np.random.default_rng(seed)
    d = O.BrainDims(N, stages=S, per_slot=False)
    L = O.init_params(d, rng, bias_scale=0.05)
    for l in L:
        l["W"], l["b"] = l["W"].astype(np.float32).astype(np.float64), l["b"].astype(np.float32).astype(np.float64)
    node, edge, adj, _ = O.synth_batch(B, N, rng)
    node, edge = node.astype(np.float32), edge.astype(np.float32)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False)
    brain.set_flat_params(O.flatten_params(L), 0)
    qr = O.brain_forward(d, L, node.astype(np.float64), edge.astype(np.float64), adj)
    return brain, node, edge, adj, qr


@pytest.mark.parametrize("N,S,B", [(20, 2, 6), (20, 2, 64), (20, 2, 1000), (20, 3, 333), (4, 3, 257), (7, 1, 50),
                                   (32, 2, 77), (2, 2, 9), (20, 2, 8192)])
def test_tc_forward_matches_oracle(v2v, N, S, B):
    brain, node, edge, adj, qr = _setup(v2v, N, S, B, seed=100 + N + S)
    info = brain.tensor_core_info()
    assert info["capable"] == 1 and info["graphs_per_tile"] == 128 // N
    x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
    brain.set_tensor_core(0)
    q_fp32 = np.stack(brain.predict(x), 1)
    brain.set_tensor_core(2)
    q_tc = np.stack(brain.predict(x), 1)
    scale = np.abs(qr).max()
    err_tc, err_fp = np.abs(q_tc - qr).max() / scale, np.abs(q_fp32 - qr).max() / scale
    assert err_tc <= RTOL, (err_tc, err_fp)
    assert err_tc <= 1e-5 + 4 * err_fp, (err_tc, err_fp)          # fp32-grade, not single-pass TF32 (~1e-3)
    # target network, device entry point, and the automatic mode at a large batch
    brain.update_target_model()
    dev = lambda a: torch.from_numpy(a).cuda()
    im, _, _ = v2v.pack_adjacency(dev(adj.astype(np.float32)))
    qd = brain.forward_device(dev(node), dev(edge), in_mask=im, target=True).cpu().numpy()
    assert np.abs(qd - qr).max() / scale <= RTOL


def test_tc_forward_repeatable_and_per_slot_rejected(v2v):
    brain, node, edge, adj, _ = _setup(v2v, 20, 2, 500, seed=7)
    brain.set_tensor_core(2)
    x = {"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}
    a = np.stack(brain.predict(x), 1)
    b = np.stack(brain.predict(x), 1)
    assert np.array_equal(a, b)                                   # deterministic: fixed MMA order
    ps = v2v.BS(4, 3, 1, 16, 1, 4, data_parallel=False, seed=1)   # reference default: per-slot weights
    assert ps.tensor_core_info()["capable"] == 0
    with pytest.raises(ValueError):
        ps.set_tensor_core(2)
    ps.set_tensor_core(0)
