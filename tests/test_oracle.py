"""CPU tests of the oracle itself: literal (reference-form) vs packed forward, the manual
backward vs torch autograd, hand-derivable known answers, and the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases
from oracle import v2v_oracle as O
from oracle import torch_ref as T


def _case(N, per_slot, S, B=5, seed=7, sparse=None):
    rng = np.random.default_rng(seed)
    d = O.BrainDims(num_d2d=N, stages=S, per_slot=per_slot)
    L = O.init_params(d, rng, bias_scale=0.1)
    node, edge, adj, dest = O.synth_batch(B, N, rng, sparse_in_degree=sparse)
    return d, L, node, edge, adj, dest, rng


def test_dims_match_reference_formulas():
    d = O.BrainDims(4, 3, 1, 16, 1, 4)
    assert (d.Dn, d.De, d.D, d.num_D2D_Input) == (9, 4, 13, 68)           # BS_brain.py:101-104
    assert d.params_per_group() == 9456 and d.params_per_group() * 4 == 37824


@pytest.mark.parametrize("N,per_slot,S", [(4, True, 3), (4, False, 3), (20, False, 2), (6, True, 1), (20, True, 3)])
def test_literal_equals_packed_forward(N, per_slot, S):
    d, L, node, edge, adj, _, _ = _case(N, per_slot, S)
    A = np.stack([O.kron_adjacency(a, d.F) for a in adj])
    lit = np.stack(O.brain_forward_literal(d, L, node, edge, A), 1)
    packed = O.brain_forward(d, L, node, edge, adj)
    np.testing.assert_allclose(packed, lit, rtol=1e-12, atol=1e-10)


def test_kron_roundtrip_and_factored_aggregation():
    rng = np.random.default_rng(3)
    adj = (rng.random((3, 5, 5)) < 0.5).astype(float)
    F = 16
    A = np.stack([O.kron_adjacency(a, F) for a in adj])
    assert np.array_equal(O.adjacency_from_kron(A, F), adj)
    H = rng.normal(size=(3, 5, F))
    lit = np.stack(O.agg_layer_call([H[:, k] for k in range(5)], A), 1)
    assert np.abs(lit - O.agg_factored(H, adj)).max() == 0.0


def test_adjacency_rule_and_pairing():
    # N=4 pairing dest = [1,0,3,2]: each node aggregates exactly the other pair
    adj = O.make_adjacency([1, 0, 3, 2])
    expect = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [1, 1, 0, 0], [1, 1, 0, 0]], float)
    assert np.array_equal(adj, expect)
    # in-degree is exactly N-2 (column sums), matrix is generally asymmetric
    rng = np.random.default_rng(0)
    _, _, adj20, dest = O.synth_batch(4, 20, rng)
    assert np.all(adj20.sum(1) == 18)
    for b in range(4):
        assert np.array_equal(adj20[b], O.make_adjacency(dest[b]))


def test_known_answers_aggregation():
    H = np.zeros((1, 4, 16)); H[0, 2, 5] = 1.0                           # one-hot feature on node 2
    adj = (np.ones((4, 4)) - np.eye(4))[None]
    out = O.agg_factored(H, adj)
    assert out[0, :, 5].tolist() == [1.0, 1.0, 0.0, 1.0]                 # everyone but node 2 receives it
    assert np.all(O.agg_factored(H, np.zeros((1, 4, 4))) == 0)


def test_masks_roundtrip():
    rng = np.random.default_rng(1)
    for N in (4, 20, 32, 33, 70):
        adj = (rng.random((3, N, N)) < 0.3).astype(float)
        im, om = O.pack_masks(adj)
        W = (N + 31) // 32
        assert im.shape == (3, N, W)
        for b in range(3):
            for m in range(N):
                for n in range(N):
                    assert ((im[b, m, n // 32] >> np.uint32(n % 32)) & 1) == adj[b, n, m]
                    assert ((om[b, n, m // 32] >> np.uint32(m % 32)) & 1) == adj[b, n, m]


@pytest.mark.parametrize("N,per_slot,S,sparse", [(4, True, 3, None), (20, False, 2, None), (20, False, 3, 2), (5, True, 2, None)])
def test_manual_backward_matches_autograd(N, per_slot, S, sparse):
    d, L, node, edge, adj, _, rng = _case(N, per_slot, S, sparse=sparse)
    q = O.brain_forward(d, L, node, edge, adj)
    y = q + rng.normal(0, 1.5, q.shape)
    loss, per_head, g = O.brain_backward(d, L, node, edge, adj, y)
    tl = T.to_torch_layers(L, requires_grad=True)
    A = torch.tensor(np.stack([O.kron_adjacency(a, d.F) for a in adj]))
    qt = torch.stack(T.forward_reference_form(d, tl, torch.tensor(node), torch.tensor(edge), A), 1)
    lt, ph = T.huber_total(qt, torch.tensor(y))
    lt.backward()
    assert abs(loss - lt.item()) <= 1e-10 * max(1.0, abs(loss))
    np.testing.assert_allclose(per_head, ph.detach().numpy(), rtol=1e-10)
    for i in range(len(L)):
        gw = tl[i]['W'].grad.numpy() if tl[i]['W'].grad is not None else np.zeros_like(L[i]['W'])
        scale = max(1.0, np.abs(gw).max())
        assert np.abs(g[i]['W'] - gw).max() <= 1e-10 * scale
        assert np.abs(g[i]['b'] - tl[i]['b'].grad.numpy()).max() <= 1e-10 * scale
    # the third input of stage 0 is all zeros: its weight block is dead, gradient exactly 0
    assert np.all(g[0]['W'][:, d.Dn + d.De:] == 0)


def test_stage0_independent_of_W3():
    d, L, node, edge, adj, _, rng = _case(4, True, 3)
    q = O.brain_forward(d, L, node, edge, adj)
    L[0]['W'][:, d.Dn + d.De:] += 5.0
    assert np.array_equal(q, O.brain_forward(d, L, node, edge, adj))


def test_permutation_equivariance_per_slot():
    d, L, node, edge, adj, _, rng = _case(5, True, 3)
    perm = rng.permutation(5)
    q = O.brain_forward(d, L, node, edge, adj)
    Lp = [{'W': l['W'][perm], 'b': l['b'][perm]} for l in L]
    qp = O.brain_forward(d, Lp, node[:, perm], edge[:, perm], adj[:, perm][:, :, perm])
    np.testing.assert_allclose(qp, q[:, perm], rtol=1e-10, atol=1e-10)


def test_huber_and_adam_known_values():
    e = np.array([-3.0, -1.0, -0.5, 0.0, 0.25, 1.0, 2.0])
    np.testing.assert_allclose(O.huber_elem(e), [2.5, 0.5, 0.125, 0.0, 0.03125, 0.5, 1.5])
    # first Adam step moves every coordinate by ~lr*sign(g) (bias-corrected), eps outside the sqrt
    p, m, v = O.keras_adam_step(np.zeros(3), np.array([1.0, -2.0, 0.0]), np.zeros(3), np.zeros(3), 1)
    np.testing.assert_allclose(p, [-1e-3, 1e-3, 0.0], rtol=1e-5)
    g = np.array([1e-3]); lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.5)
    p, m, v = O.keras_adam_step(np.zeros(1), g, np.zeros(1), np.zeros(1), 1)
    np.testing.assert_allclose(p, -lr_t * (0.5 * g) / (np.sqrt(0.001 * g * g) + 1e-7))


def test_td_targets_rule():
    rng = np.random.default_rng(5)
    p, pn = rng.normal(size=(6, 4, 4)), rng.normal(size=(6, 4, 4))
    a = rng.integers(0, 4, (6, 4)); r = rng.normal(size=6)
    y = O.td_targets(p, pn, a, r, 0.5)
    for b in range(6):
        for k in range(4):
            t = p[b, k].copy(); t[a[b, k]] = r[b] + 0.5 * np.amax(pn[b, k])   # BS_brain.py:684-690
            assert np.array_equal(y[b, k], t)


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = O.BrainDims(int(z["N"]), stages=int(z["S"]), per_slot=bool(z["per_slot"]))
    L = O.unflatten_params(d, z["params"].astype(np.float64))
    node, edge, adj = (z[k].astype(np.float64) for k in ("node", "edge", "adj"))
    q = O.brain_forward(d, L, node, edge, adj)
    np.testing.assert_allclose(q, z["q"], rtol=1e-12, atol=1e-12)
    loss, per_head, g = O.brain_backward(d, L, node, edge, adj, z["y"].astype(np.float64))
    assert abs(loss - float(z["loss"])) <= 1e-12 * abs(loss)
    np.testing.assert_allclose(O.flatten_params(g), z["grads"], rtol=1e-10, atol=1e-12)
    # and the fp32 torch restatement in the reference's form agrees with the fp64 oracle
    tl = T.to_torch_layers(L, dtype=torch.float32)
    A = torch.tensor(np.stack([O.kron_adjacency(a, d.F) for a in adj]), dtype=torch.float32)
    q32 = torch.stack(T.forward_reference_form(d, tl, torch.tensor(node, dtype=torch.float32),
                                               torch.tensor(edge, dtype=torch.float32), A), 1).numpy()
    assert np.abs(q32 - z["q"]).max() <= 1e-5 * np.abs(z["q"]).max()


def test_golden_env_states_follow_reference_rules():
    z = np.load(os.path.join(GOLDEN, "ref_n4_env.npz"))
    assert z["node"].shape[1:] == (4, 9) and z["edge"].shape[1:] == (4, 4)
    assert np.all(z["node"][:, :, 8] == 10.0)             # fixed V2V power 10 dBm (Environment.py:194-195)
    assert np.all(z["adj"].sum(1) == 2)                   # in-degree N-2
    assert np.all(np.diagonal(z["adj"], axis1=1, axis2=2) == 0)


def test_huber_and_adam_against_independent_torch_implementations():
    """Cross-checks against code that was not written for this repo (torch's Huber loss and Adam): the oracle's Huber
    element/mean reduction is torch.nn.functional.huber_loss, and its Keras-Adam rule coincides with torch.optim.Adam when
    eps -> 0 (the two differ only in where eps enters: Keras 2.2.4 adds it to sqrt(v) AFTER folding the bias corrections
    into lr_t, BS_brain.py:212)."""
    import torch
    rng = np.random.default_rng(11)
    q, y = rng.normal(0, 2, (32, 5, 4)), rng.normal(0, 2, (32, 5, 4))
    total, per_head = O.brain_loss(q, y)
    ref = [float(torch.nn.functional.huber_loss(torch.from_numpy(q[:, k]), torch.from_numpy(y[:, k]), delta=1.0, reduction="mean"))
           for k in range(5)]
    np.testing.assert_allclose(per_head, ref, rtol=1e-12)
    assert abs(total - sum(ref)) <= 1e-12 * abs(total)
    # Adam: 6 steps on a random quadratic, eps = 1e-30 in both rules, fp64
    p0 = rng.normal(size=50)
    pt = torch.tensor(p0.copy(), dtype=torch.float64, requires_grad=True)
    b1, b2, lr = float(np.float32(0.5)), float(np.float32(0.999)), float(np.float32(1e-3))
    opt = torch.optim.Adam([pt], lr=lr, betas=(b1, b2), eps=1e-30)
    p, m, v = p0.copy(), np.zeros(50), np.zeros(50)
    A = rng.normal(size=(50, 50)); A = A @ A.T / 50 + np.eye(50)
    for t in range(1, 7):
        g = A @ p
        p, m, v = O.keras_adam_step(p, g, m, v, t, eps=1e-30)
        opt.zero_grad()
        (0.5 * pt @ torch.from_numpy(A) @ pt).backward()
        opt.step()
        np.testing.assert_allclose(p, pt.detach().numpy(), rtol=1e-9, atol=1e-12)


def test_known_answers_literal_kronecker_form_and_identity_weights():
    """Hand-derivable cases in the reference's LITERAL form: AggLayer.call as concat -> batch_dot against
    kron(Adj, I_F) (BS_brain.py:69-76, :492-493) and GNNLayer.call (:44-51) with identity-like weights."""
    N, F = 4, 16
    # one-hot feature on node 2, Adj = 1 - I with node 0's own receiver (node 1) cleared as well (BS_brain.py:441-445)
    adj = np.ones((N, N)) - np.eye(N)
    adj[1, 0] = 0.0
    A = O.kron_adjacency(adj, F)[None]
    D = [np.zeros((1, F)) for _ in range(N)]
    D[2][0, 5] = 1.0
    out = O.agg_layer_call(D, A)
    assert [o[0, 5] for o in out] == [1.0, 1.0, 0.0, 1.0] and sum(np.abs(o).sum() for o in out) == 3.0
    D = [np.full((1, F), float(k + 1)) for k in range(N)]               # node k carries the constant k + 1
    out = O.agg_layer_call(D, A)
    assert [o[0, 0] for o in out] == [3.0 + 4.0, 1.0 + 3.0 + 4.0, 1.0 + 2.0 + 4.0, 1.0 + 2.0 + 3.0]   # column sums of adj
    # GNNLayer with W1 = [I; 0], W2 = 0, W3 = I, bias = 1: out = a[:, :F] + c + 1 (linear) and relu of it
    a = np.arange(25, dtype=np.float64)[None] - 12.0
    b = np.ones((1, 4)); c = np.full((1, F), 0.5)
    W1 = np.concatenate([np.eye(F), np.zeros((9, F))], 0)
    lin = O.gnn_layer_call(a, b, c, W1, np.zeros((4, F)), np.eye(F), np.ones(F))
    assert np.array_equal(lin, a[:, :F] + 1.5)
    assert np.array_equal(O.gnn_layer_call(a, b, c, W1, np.zeros((4, F)), np.eye(F), np.ones(F), 'relu'), np.maximum(a[:, :F] + 1.5, 0))
    # whole brain, all weights zero: Q = bias of the output layer; output weights = 1: Q = sum(relu(b3)) + b4
    d = O.BrainDims(N, stages=3, per_slot=True)
    L = [{'W': np.zeros((d.G, K, n)), 'b': np.zeros((d.G, n))} for K, n in d.layer_shapes()]
    L[-1]['b'][:] = np.arange(4.0)
    rng = np.random.default_rng(0)
    node, edge, adjb, _ = O.synth_batch(3, N, rng)
    assert np.array_equal(O.brain_forward(d, L, node, edge, adjb), np.broadcast_to(np.arange(4.0), (3, N, 4)))
    L[-2]['b'][:] = np.linspace(-1.0, 1.0, 20)
    L[-1]['W'][:] = 1.0
    want = np.maximum(np.linspace(-1.0, 1.0, 20), 0).sum() + np.arange(4.0)
    assert np.allclose(O.brain_forward(d, L, node, edge, adjb), want, rtol=0, atol=1e-12)
    lit = O.brain_forward_literal(d, L, node, edge, O.kron_adjacency(adjb, F))
    assert np.allclose(np.stack(lit, 1), want, rtol=0, atol=1e-12)
