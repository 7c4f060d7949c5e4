"""CPU oracle for the V2V graph-convolution hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy restatement of the arithmetic the reference performs in
``/root/reference/BS_brain.py:17-239`` through Keras 2.2.4 / TensorFlow 1.14.0
(``README.md:9-11``; neither is vendored in the reference tree nor installable
here).  It is the *checker* for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package never does.

PARITY: the reference ships no tests, golden vectors or trained weights
 and its runtime (Keras 2.2.4 / TF 1.14) cannot run in this
image, so byte-for-byte TF1 output is NOT available: against real TensorFlow
this oracle is still "parity unpinned" (tests/golden/make_tf1_golden.py is the
committed recipe that closes it wherever that stack exists).  What it IS pinned
to, by code that runs here:
 * the reference's OWN model code -- the unmodified ``BS_brain.py`` (GNNLayer,
   AggLayer, BS._create_model, train_dnn, predict, Agent.replay,
   Agent.generate_d2d_transition) executed on tests/keras_shim, a stand-in that
   restates only the Keras/TF primitives it calls; recordings in
   tests/golden/refshim_*.npz, checks in tests/test_tf1_golden.py and
   tests/test_refshim_agent.py.  Every wiring decision (what is concatenated
   with what, which layers are shared, the TD rule as executed) is therefore
   the reference's, not a reading of it;
 * the adjacency/feature packing: the unmodified ``Environment.py``
   (tests/golden/make_golden.py);
 * the manual backward: an independent torch-autograd restatement
   (oracle/torch_ref.py) in fp64.

Every function cites the reference lines it follows.  All math is dtype-generic
(pass fp64 arrays for the high-precision oracle, fp32 for like-for-like).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# dimensions (BS_brain.py:94-104)
# ----------------------------------------------------------------------------


class BrainDims:
    """Sizes derived exactly as ``BS.__init__`` does (BS_brain.py:94-104)."""

    def __init__(self, num_d2d=4, input_node_info=3, input_edge_info=1,
                 num_d2d_feedback=16, num_d2d_neighbor=1, num_ch=4,
                 stages=3, per_slot=True, hidden=(80, 40, 20)):
        self.N = int(num_d2d)
        self.F = int(num_d2d_feedback)
        self.CH = int(num_ch)
        self.Dn = ((input_node_info - 1) * num_ch + 1) * num_d2d_neighbor   # :101
        self.De = input_edge_info * num_ch                                    # :102
        self.D = self.Dn + self.De                                            # :103
        self.num_D2D_Input = self.N * self.D + self.N ** 2                    # :104
        self.S = int(stages)
        self.per_slot = bool(per_slot)
        self.G = self.N if per_slot else 1
        self.hidden = tuple(hidden)

    def layer_shapes(self):
        """(K_total, N_out) of each layer's stacked weight, GNN stages then MLP.

        GNN stage s stacks ``[W1; W2; W3]`` (BS_brain.py:26-37): stage 0 has
        d_a = Dn (``:147``), later stages d_a = F + Dn (``:154-164``).
        Decision MLP input is ``[node | h | agg]`` = Dn + 2F (``:168-175``).
        """
        shapes = []
        for s in range(self.S):
            da = self.Dn if s == 0 else self.F + self.Dn
            shapes.append((da + self.De + self.F, self.F))
        k = self.Dn + 2 * self.F
        for h in self.hidden:
            shapes.append((k, h))
            k = h
        shapes.append((k, self.CH))
        return shapes

    def params_per_group(self):
        return sum(k * n + n for k, n in self.layer_shapes())


# ----------------------------------------------------------------------------
# initialisers (Keras 'glorot_uniform', 'zeros'; BS_brain.py:26-41, Dense :176-200)
# ----------------------------------------------------------------------------


def glorot_uniform(rng, fan_in, fan_out, dtype=np.float64):
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=(fan_in, fan_out)).astype(dtype)


def init_params(dims: BrainDims, rng, dtype=np.float64, bias_scale=0.0):
    """Per layer ``{'W': [G, K, Nout], 'b': [G, Nout]}``.

    GNN layers draw W1, W2, W3 independently with their own fan-in, as three
    ``add_weight`` calls do (BS_brain.py:26-37), then stack them row-wise in
    the order [W1; W2; W3].  ``bias_scale`` > 0 gives non-zero biases so that
    parity tests exercise the bias path (Keras initialises zeros, ``:38-41``).
    """
    layers = []
    G, F, Dn, De = dims.G, dims.F, dims.Dn, dims.De
    for li, (K, Nout) in enumerate(dims.layer_shapes()):
        W = np.empty((G, K, Nout), dtype)
        for g in range(G):
            if li < dims.S:
                da = K - De - F
                W[g] = np.concatenate([glorot_uniform(rng, da, Nout, dtype),
                                       glorot_uniform(rng, De, Nout, dtype),
                                       glorot_uniform(rng, F, Nout, dtype)], 0)
            else:
                W[g] = glorot_uniform(rng, K, Nout, dtype)
        b = (bias_scale * rng.standard_normal((G, Nout))).astype(dtype)
        layers.append({'W': W, 'b': b})
    return layers


def flatten_params(layers):
    """Flat engine layout: per layer, W [G,K,Nout] then b [G,Nout]."""
    return np.concatenate([np.concatenate([l['W'].ravel(), l['b'].ravel()]) for l in layers])


def unflatten_params(dims: BrainDims, flat):
    layers, o = [], 0
    for K, Nout in dims.layer_shapes():
        nW, nb = dims.G * K * Nout, dims.G * Nout
        W = flat[o:o + nW].reshape(dims.G, K, Nout); o += nW
        b = flat[o:o + nb].reshape(dims.G, Nout); o += nb
        layers.append({'W': W.copy(), 'b': b.copy()})
    assert o == flat.size
    return layers


# ----------------------------------------------------------------------------
# adjacency (BS_brain.py:441-445, :492-493, :603)
# ----------------------------------------------------------------------------


def make_adjacency(dest):
    """``Adj = 1 - I``; ``Adj[n, m] = 0`` where ``n == dest[m]`` (:441-445)."""
    dest = np.asarray(dest)
    N = dest.shape[-1]
    adj = np.ones((N, N)) - np.eye(N)
    for n in range(N):
        for m in range(N):
            if n == dest[m]:
                adj[n, m] = 0
    return adj


def kron_adjacency(adj, F):
    """Dense Kronecker operand the reference feeds to AggLayer (:492-493, :603)."""
    return np.kron(adj, np.eye(F))


def adjacency_from_kron(A, F):
    """Exact inverse of kron_adjacency (strided sampling)."""
    return np.ascontiguousarray(A[..., ::F, ::F])


def pack_masks(adj):
    """adj [B,N,N] (0/1) -> (in_mask, out_mask) uint32 [B,N,W], W = ceil(N/32).

    in_mask[b,m] bit n  = adj[b,n,m]  (forward: who m gathers from)
    out_mask[b,n] bit m = adj[b,n,m]  (backward: who n scatters to)
    """
    adj = np.asarray(adj)
    B, N, _ = adj.shape
    W = (N + 31) // 32
    in_mask = np.zeros((B, N, W), np.uint32)
    out_mask = np.zeros((B, N, W), np.uint32)
    nz = adj != 0
    for j in range(N):
        w, bit = j // 32, np.uint32(1 << (j % 32))
        in_mask[:, :, w] |= np.where(nz[:, j, :], bit, np.uint32(0))     # n = j
        out_mask[:, :, w] |= np.where(nz[:, :, j], bit, np.uint32(0))    # m = j
    return in_mask, out_mask


# ----------------------------------------------------------------------------
# the two custom layers, literal form
# ----------------------------------------------------------------------------


def relu(x):
    return np.maximum(x, 0)


def gnn_layer_call(a, b, c, W1, W2, W3, bias, activation=None):
    """GNNLayer.call (BS_brain.py:44-51): act(a.W1 + b.W2 + c.W3 + bias)."""
    out = a @ W1 + b @ W2 + c @ W3
    out = out + bias
    if activation == 'relu':
        out = relu(out)
    elif activation is not None:
        raise ValueError(activation)
    return out


def agg_layer_call(D_list, A):
    """AggLayer.call (BS_brain.py:69-76), generalised from 4 to len(D_list) slots.

    ``K.concatenate(axis=-1)`` then ``K.batch_dot(D, A, axes=[1, 1])`` which for
    a 2-D x 3-D pair is out[b, j] = sum_i D[b, i] * A[b, i, j]; then sliced into
    per-slot (B, F) blocks.
    """
    F = D_list[0].shape[1]
    D = np.concatenate(D_list, axis=-1)
    out = np.einsum('bi,bij->bj', D, A)
    return [out[:, k * F:(k + 1) * F] for k in range(len(D_list))]


def agg_factored(H, adj):
    """Factored aggregation proven equal to agg_layer_call with A = kron(adj, I):
    agg[b, m, :] = sum_n adj[b, n, m] * H[b, n, :]."""
    return np.einsum('bnm,bnf->bmf', adj, H)


def agg_factored_T(dAgg, adj):
    """Backward of agg_factored w.r.t. H: dH[b,n,:] = sum_m adj[b,n,m] dAgg[b,m,:]."""
    return np.einsum('bnm,bmf->bnf', adj, dAgg)


# ----------------------------------------------------------------------------
# brain forward: literal per-slot wiring (BS_brain.py:108-208)
# ----------------------------------------------------------------------------


def _split_gnn(dims, W, s):
    da = dims.Dn if s == 0 else dims.F + dims.Dn
    return W[:da], W[da:da + dims.De], W[da + dims.De:]


def brain_forward_literal(dims: BrainDims, layers, node, edge, A_kron, neigh=None):
    """Reference-form forward: per-slot layer objects, Kronecker aggregation.

    node [B,N,Dn], edge [B,N,De], A_kron [B,N*F,N*F].  Returns list of N (B,CH)
    arrays like ``Model.predict`` (:208, :229-231).  Stage 0's third input is
    zeros (:478, :589); the last GNN stage is linear (:161-164).
    """
    B, N = node.shape[:2]
    S = dims.S
    g = (lambda k: k) if dims.per_slot else (lambda k: 0)
    zeros = np.zeros((B, dims.F), node.dtype)
    D = []
    for k in range(N):
        W1, W2, W3 = _split_gnn(dims, layers[0]['W'][g(k)], 0)
        act = 'relu' if S > 1 else None
        c0 = zeros if neigh is None else neigh[:, k]
        D.append(gnn_layer_call(node[:, k], edge[:, k], c0, W1, W2, W3, layers[0]['b'][g(k)], act))
    Agg = agg_layer_call(D, A_kron)
    for s in range(1, S):
        act = 'relu' if s < S - 1 else None
        Dnew = []
        for k in range(N):
            W1, W2, W3 = _split_gnn(dims, layers[s]['W'][g(k)], s)
            a = np.concatenate([D[k], node[:, k]], axis=-1)                      # :154
            Dnew.append(gnn_layer_call(a, edge[:, k], Agg[k], W1, W2, W3, layers[s]['b'][g(k)], act))
        D = Dnew
        Agg = agg_layer_call(D, A_kron)
    outs = []
    for k in range(N):
        x = np.concatenate([node[:, k], np.concatenate([D[k], Agg[k]], -1)], -1)  # :168-175
        nl = len(layers) - S
        for j in range(nl):
            L = layers[S + j]
            x = x @ L['W'][g(k)] + L['b'][g(k)]
            if j < nl - 1:
                x = relu(x)                                                      # :176-179
        outs.append(x)
    return outs


# ----------------------------------------------------------------------------
# brain forward/backward: packed form (what the CUDA engine computes)
# ----------------------------------------------------------------------------


def _gmm(x, W):
    """x [B,N,K], W [G,K,O] with G in {1,N} -> [B,N,O]."""
    if W.shape[0] == 1:
        return x @ W[0]
    return np.einsum('bnk,nko->bno', x, W)


def brain_forward(dims: BrainDims, layers, node, edge, adj, keep=False, neigh=None):
    """Packed forward.  Returns Q [B,N,CH] (and the tape when ``keep``).  ``neigh`` is the
    D{k}_Neighbor_Input of stage 0, which the reference always feeds as zeros (:478, :589)."""
    S = dims.S
    tape = {'x': [], 'pre': []}
    x = np.concatenate([node, edge] + ([] if neigh is None else [neigh]), -1)   # stage 0 active inputs
    K0 = x.shape[-1]
    pre = _gmm(x, layers[0]['W'][:, :K0]) + layers[0]['b'][None]
    h = relu(pre) if S > 1 else pre
    tape['x'].append(x); tape['pre'].append(pre)
    agg = agg_factored(h, adj)
    for s in range(1, S):
        x = np.concatenate([h, node, edge, agg], -1)
        pre = _gmm(x, layers[s]['W']) + layers[s]['b'][None]
        h = relu(pre) if s < S - 1 else pre
        tape['x'].append(x); tape['pre'].append(pre)
        agg = agg_factored(h, adj)
    x = np.concatenate([node, h, agg], -1)
    nl = len(layers) - S
    for j in range(nl):
        L = layers[S + j]
        pre = _gmm(x, L['W']) + L['b'][None]
        tape['x'].append(x); tape['pre'].append(pre)
        x = relu(pre) if j < nl - 1 else pre
    if keep:
        return x, tape
    return x


def huber_elem(e, delta=1.0):
    """tf.losses.huber_loss element (BS_brain.py:86-87)."""
    ae = np.abs(e)
    quad = np.minimum(ae, delta)
    lin = ae - quad
    return 0.5 * quad * quad + delta * lin


def brain_loss(q, y):
    """Per-head mean over (B, CH) -- SUM_BY_NONZERO_WEIGHTS -- heads summed
    (loss_weights = 1; BS_brain.py:214).  q, y [B,N,CH].  Returns (total, per_head[N])."""
    per_head = huber_elem(q - y).mean(axis=(0, 2))
    return per_head.sum(), per_head


def brain_backward(dims: BrainDims, layers, node, edge, adj, y, neigh=None, q_for_loss=None):
    """Manual reverse pass of brain_forward + brain_loss.

    Returns (loss, per_head, grads) with grads in the same structure as layers.
    ``q_for_loss`` (optional) replaces the network output inside the Huber residual q - y only:
    tests use it to feed the fp32-rounded output of the device, which separates the parity of the
    backward kernels from the cancellation in q - y when |q| >> |q - y|.
    """
    S = dims.S
    F, Dn, De = dims.F, dims.Dn, dims.De
    q, tape = brain_forward(dims, layers, node, edge, adj, keep=True, neigh=neigh)
    B = q.shape[0]
    if q_for_loss is not None:
        q = np.asarray(q_for_loss, dtype=q.dtype)
    loss, per_head = brain_loss(q, y)
    grads = [{'W': np.zeros_like(l['W']), 'b': np.zeros_like(l['b'])} for l in layers]

    def wgrad(li, x, dz):
        Kx = x.shape[-1]
        if dims.G == 1:
            grads[li]['W'][0, :Kx] = np.einsum('bnk,bno->ko', x, dz)
            grads[li]['b'][0] = dz.sum((0, 1))
        else:
            grads[li]['W'][:, :Kx] = np.einsum('bnk,bno->nko', x, dz)
            grads[li]['b'][:] = dz.sum(0)

    def dgrad(li, dz):
        W = layers[li]['W']
        if dims.G == 1:
            return dz @ W[0].T
        return np.einsum('bno,nko->bnk', dz, W)

    dz = np.clip(q - y, -1.0, 1.0) / (B * dims.CH)
    nl = len(layers) - S
    for j in reversed(range(nl)):
        li = S + j
        if j < nl - 1:
            dz = dz * (tape['pre'][li] > 0)
        wgrad(li, tape['x'][li], dz)
        dz = dgrad(li, dz)
    # dz is now d[node | h | agg]
    dh = dz[..., Dn:Dn + F] + agg_factored_T(dz[..., Dn + F:], adj)
    for s in reversed(range(1, S)):
        dpre = dh * (tape['pre'][s] > 0) if s < S - 1 else dh
        wgrad(s, tape['x'][s], dpre)
        dx = dgrad(s, dpre)                                 # d[h | node | edge | agg]
        dh = dx[..., :F] + agg_factored_T(dx[..., F + Dn + De:], adj)
    dpre = dh * (tape['pre'][0] > 0) if S > 1 else dh
    wgrad(0, tape['x'][0], dpre)                            # W3 rows of stage 0 stay 0
    return loss, per_head, grads


# ----------------------------------------------------------------------------
# optimiser: Keras 2.2.4 Adam (BS_brain.py:212), epsilon = K.epsilon() = 1e-7
# ----------------------------------------------------------------------------


def keras_adam_step(p, g, m, v, t, lr=1e-3, beta1=0.5, beta2=0.999, eps=1e-7):
    """One update; ``t`` is the 1-based iteration.  Returns (p, m, v).

    Keras keeps lr / beta_1 / beta_2 as float32 ``K.variable``s and evaluates the rule in
    float32, so the hyper-parameters are rounded to fp32 first (0.999 -> 0.99900001287...,
    which changes 1 - beta_2 by 1.3e-5 relative); the arithmetic itself stays in the dtype of
    the arrays passed in (fp64 for the high-precision oracle).
    """
    lr, beta1, beta2, eps = (float(np.float32(x)) for x in (lr, beta1, beta2, eps))
    lr_t = lr * (np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


# ----------------------------------------------------------------------------
# DQN target rule (BS_brain.py:668-692)
# ----------------------------------------------------------------------------


def td_targets(p, p_next, actions, rewards, gamma):
    """y = p except y[b, k, a[b,k]] = r[b] + gamma * max_a p_next[b, k, a].

    p, p_next [B,N,CH]; actions [B,N] int; rewards [B].
    """
    y = np.array(p, copy=True)
    B, N, _ = y.shape
    tgt = rewards[:, None] + gamma * p_next.max(-1)
    bi, ki = np.meshgrid(np.arange(B), np.arange(N), indexing='ij')
    y[bi, ki, actions] = tgt
    return y


# ----------------------------------------------------------------------------
# synthetic inputs with the distributions of SURVEY.md 8(d)
# ----------------------------------------------------------------------------


def synth_batch(B, N, rng, CH=4, dtype=np.float64, sparse_in_degree=None):
    """Node [B,N,2CH+1], edge [B,N,CH], adj [B,N,N], dest [B,N].

    Dense variant: adjacency per make_adjacency with dest(m) uniform over the
    other nodes (E = N(N-2)).  Sparse variant: ``sparse_in_degree`` random
    in-neighbours per node (E = N * in_degree).
    """
    big = N > 8
    v2v = rng.normal(0.66 if big else 0.93, 0.44 if big else 0.15, (B, N, CH))
    v2i = rng.normal(0.54 if big else 0.72, 0.17 if big else 0.20, (B, N, CH))
    pwr = np.full((B, N, 1), 10.0)
    node = np.concatenate([v2v, v2i, pwr], -1).astype(dtype)
    edge = rng.normal(0.92, 0.11, (B, N, CH)).astype(dtype)
    dest = (np.arange(N)[None, :] + rng.integers(1, N, (B, N))) % N
    if sparse_in_degree is None:
        adj = np.ones((B, N, N)) - np.eye(N)[None]
        b_idx = np.repeat(np.arange(B), N)
        adj[b_idx, dest.ravel(), np.tile(np.arange(N), B)] = 0.0
    else:
        adj = np.zeros((B, N, N))
        for k in range(sparse_in_degree):
            src = (np.arange(N)[None, :] + 1 + (rng.integers(0, N - 1, (B, N)) + k) % (N - 1)) % N
            b_idx = np.repeat(np.arange(B), N)
            adj[b_idx, src.ravel(), np.tile(np.arange(N), B)] = 1.0
    return node, edge, adj.astype(dtype), dest
