"""Host-side mirror of the reference's two custom Keras layers.

``GNNLayer`` (BS_brain.py:17-56) and ``AggLayer`` (BS_brain.py:60-82) keep the
reference's constructor / ``build`` / ``call`` / ``compute_output_shape`` surface and
its assertions; the arithmetic runs in the sm_100a kernels behind the C-ABI
(include/v2v_gnn.h).  Inputs may be numpy arrays (results come back as numpy, like
``Model.predict``) or CUDA torch tensors (results stay on the device and are
differentiable through ``torch.autograd``).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import ptr


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("the V2V GNN engine needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dtype=torch.float32):
    """numpy / torch (any device) -> contiguous CUDA tensor; returns (tensor, was_numpy)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=_device(), dtype=dtype).contiguous(), False
    arr = np.ascontiguousarray(np.asarray(x), dtype=np.float32 if dtype == torch.float32 else None)
    return torch.from_numpy(arr).to(_device(), dtype=dtype), True


def _seg_arrays(tensors):
    n = len(tensors)
    segs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    widths = (C.c_int * n)(*[int(t.shape[-1]) for t in tensors])
    return segs, widths


# ---------------------------------------------------------------------------
# raw operator wrappers (device tensors in, device tensors out)
# ---------------------------------------------------------------------------

def adjacency_from_input(A, n_nodes: int, feat: int):
    """Recover Adj [B,N,N] from what the reference feeds as ``Adjacency_Matrix``.

    The reference passes ``kron(Adj, I_F)`` of shape (B, N*F, N*F) (BS_brain.py:492-493,
    :603); sampling every F-th row/column is its exact inverse.  A raw (B, N, N)
    adjacency is accepted as is.
    """
    if A.ndim != 3:
        raise ValueError(f"Adjacency_Matrix must be 3-D, got shape {tuple(A.shape)}")
    if A.shape[1] == n_nodes and A.shape[2] == n_nodes:
        return A
    if A.shape[1] == n_nodes * feat and A.shape[2] == n_nodes * feat:
        return A[:, ::feat, ::feat]
    raise ValueError(f"Adjacency_Matrix shape {tuple(A.shape)} matches neither (B,{n_nodes},{n_nodes}) nor "
                     f"(B,{n_nodes * feat},{n_nodes * feat})")


def pack_adjacency(adj: torch.Tensor):
    """adj fp32 CUDA [B,N,N] -> (in_mask, out_mask, is_binary)."""
    lib = _lib.load()
    B, N, _ = adj.shape
    W = (N + 31) // 32
    in_mask = torch.empty((B, N, W), dtype=torch.int32, device=adj.device)
    out_mask = torch.empty((B, N, W), dtype=torch.int32, device=adj.device)
    flag = torch.zeros(1, dtype=torch.int32, device=adj.device)
    _lib.check(lib.v2v_adj_pack_masks(ptr(adj), B, N, ptr(in_mask), ptr(out_mask), ptr(flag), _lib.current_stream()),
               ValueError)
    return in_mask, out_mask, int(flag.item()) == 0


def aggregate(H: torch.Tensor, mask: torch.Tensor = None, adj: torch.Tensor = None, transpose: bool = False,
              addend: torch.Tensor = None):
    """out[b,m] = sum_n Adj[b,n,m] H[b,n] (+ addend).  ``mask`` = in_mask (or out_mask with
    the transposed meaning); ``adj`` selects the weighted kernel."""
    lib = _lib.load()
    B, N, F = H.shape
    out = torch.empty_like(H)
    if mask is not None:
        dt = {torch.float32: 0, torch.bfloat16: 1}[H.dtype]
        _lib.check(lib.v2v_agg_mask(ptr(H), ptr(mask), ptr(addend), ptr(out), B, N, F, dt, _lib.current_stream()),
                   ValueError)
    else:
        _lib.check(lib.v2v_agg_dense(ptr(H), ptr(adj), ptr(addend), ptr(out), B, N, F, int(transpose),
                                     _lib.current_stream()), ValueError)
    return out


class _AggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, in_mask, out_mask, adj):
        ctx.out_mask, ctx.adj = out_mask, adj
        return aggregate(H, in_mask, adj, transpose=False)

    @staticmethod
    def backward(ctx, g):
        return aggregate(g.contiguous(), ctx.out_mask, ctx.adj, transpose=True), None, None, None


def dense_forward(segs, W, bias, n_nodes: int, groups: int, act: bool):
    """segs: list of [B*N, w] (or [B,N,w]) fp32 CUDA tensors; W [G,K,O]; bias [G,O]."""
    lib = _lib.load()
    rows = segs[0].numel() // segs[0].shape[-1]
    B = rows // n_nodes
    O = W.shape[-1]
    out = torch.empty((rows, O), dtype=torch.float32, device=W.device)
    sp, sw = _seg_arrays(segs)
    _lib.check(lib.v2v_dense_fwd(len(segs), sp, sw, ptr(W), W.shape[-2], ptr(bias), ptr(out), B, n_nodes, groups, O,
                                 int(act), _lib.current_stream()), ValueError)
    return out


class _GNNFn(torch.autograd.Function):
    """act([a|b|c] . W + bias) with the engine's data/weight gradient kernels."""

    @staticmethod
    def forward(ctx, a, b, c, W, bias, act):
        out = dense_forward([a, b, c], W, bias, 1, 1, act)
        ctx.save_for_backward(a, b, c, W, out)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        a, b, c, W, out = ctx.saved_tensors
        g = g.contiguous()
        gate = out if ctx.act else None
        rows, O = out.shape
        K = W.shape[-2]
        st = _lib.current_stream()
        dW = torch.zeros_like(W)
        db = torch.zeros((1, O), dtype=torch.float32, device=W.device)
        sp, sw = _seg_arrays([a, b, c])
        _lib.check(lib.v2v_dense_bwd_weight(3, sp, sw, ptr(g), ptr(gate), ptr(dW), K, ptr(db), rows, 1, 1, O, st))
        grads = []
        k0 = 0
        for t, need in zip((a, b, c), ctx.needs_input_grad[:3]):
            w = t.shape[-1]
            if need:
                dx = torch.empty_like(t)
                # widths that are not a multiple of 4 take the engine's scalar kernel
                _lib.check(lib.v2v_dense_bwd_data(ptr(g), ptr(gate), ptr(W), K, k0, w, ptr(dx), 0, 0, None, None,
                                                  rows, 1, 1, O, st))
                grads.append(dx)
            else:
                grads.append(None)
            k0 += w
        return grads[0], grads[1], grads[2], dW, db.view(-1), None


# ---------------------------------------------------------------------------
# the reference's layer classes
# ---------------------------------------------------------------------------

class _LayerBase:
    _uid = 0

    def __init__(self, **kwargs):
        name = kwargs.pop("name", None)
        kwargs.pop("trainable", None)
        kwargs.pop("dtype", None)
        if kwargs:
            raise TypeError(f"Keyword argument not understood: {sorted(kwargs)[0]}")
        if name is None:
            _LayerBase._uid += 1
            name = f"{type(self).__name__.lower()}_{_LayerBase._uid}"
        self.name = name
        self.built = False

    def __call__(self, x):
        if not self.built:
            if isinstance(x, list):
                self.build([tuple([None] + list(t.shape[1:])) for t in x])
            else:
                self.build(tuple([None] + list(x.shape[1:])))
        return self.call(x)


def _activation_get(identifier):
    """keras.activations.get for the two activations the reference uses (:121-164)."""
    if identifier is None or identifier == "linear":
        return None
    if identifier == "relu":
        return "relu"
    raise ValueError(f"Unknown activation function:{identifier}")


class GNNLayer(_LayerBase):
    """``act(a.W1 + b.W2 + c.W3 + bias)`` -- BS_brain.py:17-56."""

    def __init__(self, output_dim, activation=None, **kwargs):
        self.output_dim = int(output_dim)
        self.activation = _activation_get(activation)
        super().__init__(**kwargs)

    def build(self, input_shape):
        assert isinstance(input_shape, list)
        da, db, dc = int(input_shape[0][1]), int(input_shape[1][1]), int(input_shape[2][1])
        dev = _device()
        # glorot_uniform per weight with its own fan-in (three add_weight calls, :26-37); bias zeros (:38-41)
        W = torch.empty((1, da + db + dc, self.output_dim), dtype=torch.float32, device=dev)
        off = 0
        for d in (da, db, dc):
            lim = math.sqrt(6.0 / (d + self.output_dim))
            W[0, off:off + d].uniform_(-lim, lim)
            off += d
        self._W = W.requires_grad_(True)
        self._b = torch.zeros((1, self.output_dim), dtype=torch.float32, device=dev, requires_grad=True)
        self._split = (da, db, dc)
        self.built = True

    # views named as in the reference
    @property
    def W1(self):
        return self._W[0, :self._split[0]]

    @property
    def W2(self):
        return self._W[0, self._split[0]:self._split[0] + self._split[1]]

    @property
    def W3(self):
        return self._W[0, self._split[0] + self._split[1]:]

    @property
    def b(self):
        return self._b[0]

    @property
    def trainable_weights(self):
        return [self._W, self._b]

    def get_weights(self):
        return [t.detach().cpu().numpy().copy() for t in (self.W1, self.W2, self.W3, self.b)]

    def set_weights(self, weights):
        W1, W2, W3, b = weights
        stacked = np.concatenate([np.asarray(W1), np.asarray(W2), np.asarray(W3)], axis=0).astype(np.float32)
        if stacked.shape != tuple(self._W.shape[1:]):
            raise ValueError(f"Layer weight shape {tuple(self._W.shape[1:])} not compatible with provided weight "
                             f"shape {stacked.shape}")
        with torch.no_grad():
            self._W.copy_(torch.from_numpy(stacked)[None])
            self._b.copy_(torch.from_numpy(np.asarray(b, dtype=np.float32))[None])

    def call(self, x):
        assert isinstance(x, list)
        a, b, c = x
        conv = [_to_dev(t) for t in (a, b, c)]
        (a_d, was_np), (b_d, _), (c_d, _) = conv
        if a_d.ndim != 2 or b_d.ndim != 2 or c_d.ndim != 2:
            raise ValueError("GNNLayer expects three 2-D inputs (batch, features)")
        if (a_d.shape[1], b_d.shape[1], c_d.shape[1]) != self._split:
            raise ValueError(f"GNNLayer built for input widths {self._split}, got "
                             f"{(a_d.shape[1], b_d.shape[1], c_d.shape[1])}")
        if not (a_d.shape[0] == b_d.shape[0] == c_d.shape[0]):
            raise ValueError("GNNLayer inputs must share the batch dimension")
        need_grad = torch.is_grad_enabled() and not was_np
        if need_grad:
            out = _GNNFn.apply(a_d, b_d, c_d, self._W, self._b.view(-1), self.activation == "relu")
        else:
            out = dense_forward([a_d, b_d, c_d], self._W.detach(), self._b.detach(), 1, 1, self.activation == "relu")
        return out.cpu().numpy() if was_np else out

    def compute_output_shape(self, input_shape):
        assert isinstance(input_shape, list)
        shape_a, shape_b, shape_c = input_shape
        return (shape_a[0], self.output_dim)


class AggLayer(_LayerBase):
    """Neighbour aggregation -- BS_brain.py:60-82.

    ``call([D1, ..., DN, A])``: the reference hard-codes four node slots (:71); any
    number is accepted here.  ``A`` is the reference's (B, N*F, N*F) Kronecker
    operand or the raw (B, N, N) adjacency.
    """

    def __init__(self, output_dim, **kwargs):
        self.output_dim = int(output_dim)
        super().__init__(**kwargs)

    def build(self, input_shape):
        self.built = True

    def call(self, x):
        assert isinstance(x, list)
        *Ds, A = x
        if len(Ds) < 1:
            raise ValueError("AggLayer needs at least one node input and the adjacency")
        conv = [_to_dev(t) for t in Ds]
        was_np = conv[0][1]
        Dd = [c[0] for c in conv]
        N, F = len(Dd), self.output_dim
        for t in Dd:
            if t.ndim != 2 or t.shape[1] != F or t.shape[0] != Dd[0].shape[0]:
                raise ValueError(f"AggLayer node inputs must all be (batch, {F})")
        A_d, _ = _to_dev(A)
        adj = adjacency_from_input(A_d, N, F).contiguous()
        if adj.shape[0] != Dd[0].shape[0]:
            raise ValueError("Adjacency_Matrix batch dimension differs from the node inputs")
        H = torch.stack(Dd, dim=1)                      # (B, N, F): K.concatenate of :72, as a tensor
        in_mask, out_mask, binary = pack_adjacency(adj)
        if binary:
            args = (in_mask, out_mask, None)
        else:
            args = (None, None, adj)
        if torch.is_grad_enabled() and H.requires_grad:
            out = _AggFn.apply(H, *args)
        else:
            out = aggregate(H, args[0], args[2])
        outs = [out[:, k, :] for k in range(N)]          # the slices of :75-76
        if was_np:
            return [o.cpu().numpy() for o in outs]
        return outs

    def compute_output_shape(self, input_shape):
        assert isinstance(input_shape, list)
        shape_A = input_shape[-1]
        return [(shape_A[0], self.output_dim) for _ in input_shape[:-1]]
