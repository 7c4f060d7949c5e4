"""bf16-storage restatement of the packed brain (TEST INFRASTRUCTURE ONLY -- see oracle/v2v_oracle.py's header).

BASELINE.json configs[2] names a "3-layer GNN bf16" variant of the reference's brain (BS_brain.py:108-216).  The
reference itself only ever runs fp32 (Keras default), so this configuration has no reference of its own: the fp64
oracle of oracle/v2v_oracle.py stays the statement of the arithmetic, and this file restates the SAME arithmetic with
the rounding points of the engine's bf16 kernel (csrc/tc_train.cu) made explicit, so that the kernel can be checked
tightly (accumulation order and rounding-boundary flips only) in addition to the loose check against fp64:

  * operands of every contraction are bf16 (round-to-nearest-even): weights, inputs, activations, back-propagated
    gradients; products are exact, accumulation is fp32 on the device (fp64 here);
  * the neighbour aggregation (BS_brain.py:69-76) is one of those contractions (0/1 adjacency times bf16 rows): it sums
    the ROUNDED h rows (forward) and the ROUNDED dagg rows (backward, accumulated onto the unrounded dh);
  * bias add, ReLU, the Huber head (:86-87) and the output layer are fp32;
  * the weight gradient contracts the rounded activations with the rounded dz; the bias gradient sums the rounded dz;
  * master weights, gradients and Keras-Adam (:212) stay fp32 (not restated here: oracle/v2v_oracle.keras_adam_step).

PARITY UNPINNED for the same reason as v2v_oracle.py.
"""
from __future__ import annotations

import numpy as np

from . import v2v_oracle as O


def bf16(x):
    """Round to the nearest bfloat16 (ties to even), returned as float64."""
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = a.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).astype(np.float64).reshape(a.shape)


def f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def brain_forward_backward_bf16(dims: O.BrainDims, layers, node, edge, adj, y=None, q_for_loss=None):
    """Shared-weight brain (G == 1) with the engine's bf16 rounding points.

    Returns q [B,N,CH] when ``y`` is None, else (q, loss, per_head, grads) with grads structured like ``layers``.
    ``q_for_loss`` replaces the output inside the Huber residual only (as in v2v_oracle.brain_backward).
    """
    assert dims.G == 1, "the bf16 kernel is the shared-weight brain"
    S, F, Dn, De = dims.S, dims.F, dims.Dn, dims.De
    Wb = [bf16(l['W'][0]) for l in layers]
    bias = [f32(l['b'][0]) for l in layers]
    nb, eb = bf16(node), bf16(edge)
    xs, pres, hs = [], [], []
    x = np.concatenate([nb, eb], -1)
    pre = x @ Wb[0][:Dn + De] + bias[0]
    h = O.relu(pre) if S > 1 else pre
    xs.append(x); pres.append(pre)
    agg = O.agg_factored(bf16(h), adj)
    for s in range(1, S):
        x = np.concatenate([bf16(h), nb, eb, bf16(agg)], -1)
        pre = x @ Wb[s] + bias[s]
        h = O.relu(pre) if s < S - 1 else pre
        xs.append(x); pres.append(pre)
        agg = O.agg_factored(bf16(h), adj)
    x = np.concatenate([nb, bf16(h), bf16(agg)], -1)
    nl = len(layers) - S
    for j in range(nl):
        pre = x @ Wb[S + j] + bias[S + j]
        xs.append(x); pres.append(pre)
        x = bf16(O.relu(pre)) if j < nl - 1 else pre
    q = x
    if y is None:
        return q
    B = q.shape[0]
    q_out = q
    if q_for_loss is not None:
        q = np.asarray(q_for_loss, np.float64)
    loss, per_head = O.brain_loss(f32(q), y)
    grads = [{'W': np.zeros_like(np.asarray(l['W'], np.float64)), 'b': np.zeros_like(np.asarray(l['b'], np.float64))}
             for l in layers]

    def wgrad(li, xin, dzb):
        Kx = xin.shape[-1]
        grads[li]['W'][0, :Kx] = np.einsum('bnk,bno->ko', xin, dzb)
        grads[li]['b'][0] = dzb.sum((0, 1))

    dz = bf16(np.clip(f32(q) - y, -1.0, 1.0) / (B * dims.CH))
    for j in reversed(range(nl)):
        li = S + j
        wgrad(li, xs[li], dz)
        dx = dz @ Wb[li].T
        if j > 0:
            dz = bf16(dx * (pres[li - 1] > 0))               # relu of the producing MLP layer
    # dx = d[node | h | agg] of the first MLP layer
    dh = f32(dx[..., Dn:Dn + F]) + O.agg_factored_T(bf16(dx[..., Dn + F:]), adj)
    for s in reversed(range(1, S)):
        dpre = bf16(dh * (pres[s] > 0) if s < S - 1 else dh)
        wgrad(s, xs[s], dpre)
        dx = dpre @ Wb[s].T                                  # d[h | node | edge | agg]
        dh = f32(dx[..., :F]) + O.agg_factored_T(bf16(dx[..., F + Dn + De:]), adj)
    dpre = bf16(dh * (pres[0] > 0) if S > 1 else dh)
    wgrad(0, xs[0], dpre)
    return q_out, loss, per_head, grads
