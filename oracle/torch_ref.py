"""Second, independent CPU restatement in torch (TEST INFRASTRUCTURE ONLY).

Same role and import rules as ``oracle/v2v_oracle.py`` (see its header;
PARITY UNPINNED by the reference).  Two uses:

* autograd cross-check of the NumPy oracle's manual backward (tests);
* the timed CPU baseline in the reference's own *form*: one small layer object
  per node slot, ``concat -> (B,NF) x (B,NF,NF)`` ``bmm`` against the dense
  Kronecker adjacency, Keras-rule Adam (BS_brain.py:44-51, :69-76, :108-216).
"""
from __future__ import annotations

import torch


def _gnn_call(a, b, c, W1, W2, W3, bias, act):
    # BS_brain.py:47-50
    out = a @ W1 + b @ W2 + c @ W3
    out = out + bias
    return torch.relu(out) if act else out


def _agg_call(D_list, A):
    # BS_brain.py:71-76 ; K.batch_dot(D, A, axes=[1,1])
    F = D_list[0].shape[1]
    D = torch.cat(D_list, dim=-1)
    out = torch.bmm(D.unsqueeze(1), A).squeeze(1)
    return [out[:, k * F:(k + 1) * F] for k in range(len(D_list))]


def split_layers(dims, layers):
    """layers: list of {'W': [G,K,O] tensor, 'b': [G,O] tensor}."""
    return layers


def forward_reference_form(dims, layers, node, edge, A_kron):
    """Literal per-slot forward (BS_brain.py:117-200). Returns list of N (B,CH)."""
    B, N = node.shape[:2]
    S, F, Dn, De = dims.S, dims.F, dims.Dn, dims.De
    g = (lambda k: k) if dims.per_slot else (lambda k: 0)

    def split(W, s):
        da = Dn if s == 0 else F + Dn
        return W[:da], W[da:da + De], W[da + De:]

    zeros = torch.zeros(B, F, dtype=node.dtype)
    D = []
    for k in range(N):
        W1, W2, W3 = split(layers[0]['W'][g(k)], 0)
        D.append(_gnn_call(node[:, k], edge[:, k], zeros, W1, W2, W3, layers[0]['b'][g(k)], S > 1))
    Agg = _agg_call(D, A_kron)
    for s in range(1, S):
        Dn_ = []
        for k in range(N):
            W1, W2, W3 = split(layers[s]['W'][g(k)], s)
            a = torch.cat([D[k], node[:, k]], dim=-1)
            Dn_.append(_gnn_call(a, edge[:, k], Agg[k], W1, W2, W3, layers[s]['b'][g(k)], s < S - 1))
        D = Dn_
        Agg = _agg_call(D, A_kron)
    outs = []
    nl = len(layers) - S
    for k in range(N):
        x = torch.cat([node[:, k], torch.cat([D[k], Agg[k]], -1)], -1)
        for j in range(nl):
            L = layers[S + j]
            x = x @ L['W'][g(k)] + L['b'][g(k)]
            if j < nl - 1:
                x = torch.relu(x)
        outs.append(x)
    return outs


def forward_factored(dims, layers, node, edge, adj):
    """Packed forward with the factored aggregation. Returns Q [B,N,CH]."""
    S = dims.S

    def gmm(x, W):
        if W.shape[0] == 1:
            return x @ W[0]
        return torch.einsum('bnk,nko->bno', x, W)

    def agg(h):
        return torch.einsum('bnm,bnf->bmf', adj, h)

    x = torch.cat([node, edge], -1)
    h = gmm(x, layers[0]['W'][:, :x.shape[-1]]) + layers[0]['b'][None]
    if S > 1:
        h = torch.relu(h)
    a = agg(h)
    for s in range(1, S):
        x = torch.cat([h, node, edge, a], -1)
        h = gmm(x, layers[s]['W']) + layers[s]['b'][None]
        if s < S - 1:
            h = torch.relu(h)
        a = agg(h)
    x = torch.cat([node, h, a], -1)
    nl = len(layers) - S
    for j in range(nl):
        L = layers[S + j]
        x = gmm(x, L['W']) + L['b'][None]
        if j < nl - 1:
            x = torch.relu(x)
    return x


def huber_total(q_list_or_tensor, y):
    """Sum over heads of mean-over-(B,CH) Huber (BS_brain.py:86-87, :214). y [B,N,CH]."""
    if isinstance(q_list_or_tensor, (list, tuple)):
        q = torch.stack(q_list_or_tensor, dim=1)
    else:
        q = q_list_or_tensor
    e = q - y
    ae = e.abs()
    quad = torch.clamp(ae, max=1.0)
    per = 0.5 * quad * quad + (ae - quad)
    per_head = per.mean(dim=(0, 2))
    return per_head.sum(), per_head


def keras_adam_(params, grads, ms, vs, t, lr=1e-3, beta1=0.5, beta2=0.999, eps=1e-7):
    """In-place Keras-2.2.4 Adam over lists of tensors (BS_brain.py:212)."""
    import numpy as _np
    lr, beta1, beta2, eps = (float(_np.float32(x)) for x in (lr, beta1, beta2, eps))   # Keras float32 variables
    lr_t = lr * ((1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t))
    with torch.no_grad():
        for p, g, m, v in zip(params, grads, ms, vs):
            m.mul_(beta1).add_(g, alpha=1.0 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
            p.sub_(lr_t * m / (v.sqrt() + eps))


def to_torch_layers(layers, dtype=torch.float64, requires_grad=False):
    out = []
    for l in layers:
        W = torch.tensor(l['W'], dtype=dtype, requires_grad=requires_grad)
        b = torch.tensor(l['b'], dtype=dtype, requires_grad=requires_grad)
        out.append({'W': W, 'b': b})
    return out


class ReferenceFormCPU:
    """The timed CPU baseline: reference-form train step / predict on host cores.

    ``fit_step`` = one fwd + Huber + autograd bwd + Keras-Adam over a B-row batch
    (model.fit(..., batch_size=B, epochs=1), BS_brain.py:218-223);
    ``predict`` runs in chunks of 32 rows (Keras default).
    """

    def __init__(self, dims, layers_np, dtype=torch.float32, form='reference'):
        self.dims = dims
        self.form = form
        self.layers = to_torch_layers(layers_np, dtype, requires_grad=True)
        self.flat = [t for l in self.layers for t in (l['W'], l['b'])]
        self.m = [torch.zeros_like(t) for t in self.flat]
        self.v = [torch.zeros_like(t) for t in self.flat]
        self.t = 0

    def _fwd(self, node, edge, adj_or_A):
        if self.form == 'reference':
            return torch.stack(forward_reference_form(self.dims, self.layers, node, edge, adj_or_A), 1)
        return forward_factored(self.dims, self.layers, node, edge, adj_or_A)

    def predict(self, node, edge, adj_or_A, chunk=32):
        with torch.no_grad():
            outs = [self._fwd(node[i:i + chunk], edge[i:i + chunk], adj_or_A[i:i + chunk])
                    for i in range(0, node.shape[0], chunk)]
        return torch.cat(outs, 0)

    def fit_step(self, node, edge, adj_or_A, y):
        for t in self.flat:
            t.grad = None
        q = self._fwd(node, edge, adj_or_A)
        loss, per_head = huber_total(q, y)
        loss.backward()
        self.t += 1
        grads = [t.grad if t.grad is not None else torch.zeros_like(t) for t in self.flat]
        keras_adam_(self.flat, grads, self.m, self.v, self.t)
        return float(loss.detach()), per_head.detach()
