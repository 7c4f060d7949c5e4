"""GPU parity of the reference's two custom layers through their Keras-like surface
(GNNLayer BS_brain.py:17-56, AggLayer BS_brain.py:60-82)."""
import numpy as np
import pytest
import torch

from oracle import v2v_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.mark.parametrize("widths,act", [((9, 4, 16), "relu"), ((25, 4, 16), "relu"), ((25, 4, 16), None)])
def test_gnn_layer_numpy_call(v2v, widths, act):
    rng = np.random.default_rng(sum(widths))
    B = 77
    a, b, c = (rng.normal(size=(B, w)) for w in widths)
    layer = v2v.GNNLayer(16, activation=act, name="D1_GNN")
    out = layer([a, b, c])                                  # builds on first call, like Keras
    assert isinstance(out, np.ndarray) and out.shape == (B, 16)
    assert layer.compute_output_shape([(None, w) for w in widths]) == (None, 16)
    W1, W2, W3, bias = layer.get_weights()
    assert W1.shape == (widths[0], 16) and np.all(bias == 0)
    for W, d in ((W1, widths[0]), (W2, widths[1]), (W3, widths[2])):     # glorot_uniform, own fan-in each (:26-37)
        assert np.abs(W).max() <= np.sqrt(6.0 / (d + 16)) + 1e-7
    ref = O.gnn_layer_call(a, b, c, W1.astype(np.float64), W2.astype(np.float64), W3.astype(np.float64), bias, act)
    assert rel_err(out, ref) <= 1e-5
    new = [rng.normal(size=w.shape).astype(np.float32) for w in (W1, W2, W3, bias)]
    layer.set_weights(new)
    ref = O.gnn_layer_call(a, b, c, *[w.astype(np.float64) for w in new], act)
    assert rel_err(layer.call([a, b, c]), ref) <= 1e-5
    with pytest.raises(AssertionError):
        layer.call((a, b, c))                               # `assert isinstance(x, list)` (:45)
    with pytest.raises(ValueError):
        v2v.GNNLayer(16, activation="tanh")


@pytest.mark.parametrize("N", [4, 7, 20])
def test_agg_layer_numpy_call_kron_and_raw(v2v, N):
    rng = np.random.default_rng(N)
    B, F = 31, 16
    D = [rng.normal(size=(B, F)) for _ in range(N)]
    adj = O.synth_batch(B, N, rng)[2]
    A = np.stack([O.kron_adjacency(a, F) for a in adj])
    layer = v2v.AggLayer(F, name="Aggregate")
    outs = layer(D + [A])
    ref = O.agg_layer_call(D, A)
    assert len(outs) == N
    for o, r in zip(outs, ref):
        assert o.shape == (B, F) and rel_err(o, r) <= 1e-5
    outs2 = layer(D + [adj])                                # raw (B,N,N) adjacency accepted too
    assert all(np.array_equal(o, p) for o, p in zip(outs, outs2))
    assert layer.compute_output_shape([(None, F)] * N + [(None, N * F, N * F)]) == [(None, F)] * N
    Aw = np.stack([O.kron_adjacency(a, F) for a in rng.normal(size=(B, N, N))])     # weighted adjacency
    for o, r in zip(layer(D + [Aw]), O.agg_layer_call(D, Aw)):
        assert rel_err(o, r) <= 1e-5
    with pytest.raises(ValueError):
        layer(D + [np.zeros((B, N * F + 1, N * F + 1))])


def test_layers_autograd_matches_torch(v2v):
    """A two-stage slice of the brain built from the layer objects, differentiated by torch.autograd
    through the engine's backward kernels, against the same graph in plain torch (fp64)."""
    rng = np.random.default_rng(0)
    B, N, F = 40, 4, 16
    node = [torch.tensor(rng.normal(size=(B, 9)), dtype=torch.float32, device="cuda") for _ in range(N)]
    edge = [torch.tensor(rng.normal(size=(B, 4)), dtype=torch.float32, device="cuda") for _ in range(N)]
    zeros = torch.zeros((B, F), device="cuda")
    adj = torch.tensor(O.synth_batch(B, N, rng)[2], dtype=torch.float32, device="cuda")
    g1 = [v2v.GNNLayer(F, activation="relu") for _ in range(N)]
    g2 = [v2v.GNNLayer(F) for _ in range(N)]
    agg = v2v.AggLayer(F)
    D = [g1[k]([node[k], edge[k], zeros]) for k in range(N)]
    Ag = agg(D + [adj])
    D2 = [g2[k]([torch.cat([D[k], node[k]], -1), edge[k], Ag[k]]) for k in range(N)]
    Ag2 = agg(D2 + [adj])
    loss = sum((d * d).sum() for d in D2) + sum((a * a).sum() for a in Ag2)
    loss.backward()
    # reference graph in fp64 torch
    def p64(t):
        return t.detach().double().cpu().requires_grad_(True)
    W1 = [(p64(l._W), p64(l._b)) for l in g1]
    W2 = [(p64(l._W), p64(l._b)) for l in g2]
    n64 = [t.double().cpu() for t in node]; e64 = [t.double().cpu() for t in edge]; a64 = adj.double().cpu()
    Dr = [torch.relu(torch.cat([n64[k], e64[k], torch.zeros(B, F, dtype=torch.float64)], -1) @ W1[k][0][0] + W1[k][1][0]) for k in range(N)]
    Agr = torch.einsum('bnm,bnf->bmf', a64, torch.stack(Dr, 1))
    D2r = [torch.cat([Dr[k], n64[k], e64[k], Agr[:, k]], -1) @ W2[k][0][0] + W2[k][1][0] for k in range(N)]
    Ag2r = torch.einsum('bnm,bnf->bmf', a64, torch.stack(D2r, 1))
    lr = sum((d * d).sum() for d in D2r) + (Ag2r * Ag2r).sum()
    lr.backward()
    assert abs(loss.item() - lr.item()) <= 1e-5 * abs(lr.item())
    for k in range(N):
        assert rel_err(g1[k]._W.grad.cpu().numpy(), W1[k][0].grad.numpy()) <= 1e-4
        assert rel_err(g1[k]._b.grad.cpu().numpy(), W1[k][1].grad.numpy()) <= 1e-4
        assert rel_err(g2[k]._W.grad.cpu().numpy(), W2[k][0].grad.numpy()) <= 1e-4
