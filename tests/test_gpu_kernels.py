"""GPU parity of the individual kernels, called through the C-ABI, against the NumPy oracle.

Tolerance (north_star): fp32 outputs within 1e-4 relative of the fp64 oracle; the tests use
1e-5 relative to the tensor's max magnitude for fp32 kernels (they are true-fp32 FFMA/FADD) and
state the looser bf16 storage tolerance where it applies.  Integer work (masks) is bit-exact.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import v2v_oracle as O

pytestmark = pytest.mark.gpu

RTOL_F32 = 1e-5


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def dev(x, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dtype=dtype)


def rand_adj(rng, B, N, kind):
    if kind == "dense":
        return O.synth_batch(B, N, rng)[2]
    if kind == "sparse2":
        return O.synth_batch(B, N, rng, sparse_in_degree=2)[2]
    if kind == "random":
        return (rng.random((B, N, N)) < 0.4).astype(np.float64)
    if kind == "empty":
        return np.zeros((B, N, N))
    if kind == "full":
        return np.ones((B, N, N))
    raise ValueError(kind)


# --------------------------------------------------------------------------- adjacency packing
@pytest.mark.parametrize("N", [1, 4, 20, 32, 33, 64, 100])
def test_pack_masks_bit_exact(v2v, N):
    rng = np.random.default_rng(N)
    adj = (rng.random((9, N, N)) < 0.35).astype(np.float32)
    im, om, binary = v2v.pack_adjacency(dev(adj))
    rim, rom = O.pack_masks(adj)
    assert binary
    assert np.array_equal(im.cpu().numpy().view(np.uint32), rim)
    assert np.array_equal(om.cpu().numpy().view(np.uint32), rom)
    adj[3, 0, N - 1] = 0.5
    assert not v2v.pack_adjacency(dev(adj))[2]


# --------------------------------------------------------------------------- aggregation
@pytest.mark.parametrize("B,N,kind", [
    (1, 4, "dense"), (7, 4, "random"), (256, 4, "dense"), (33, 8, "random"), (5, 7, "random"),
    (1024, 20, "dense"), (1023, 20, "sparse2"), (3, 20, "random"), (130, 19, "random"), (64, 32, "random"),
    (9, 31, "full"), (50, 20, "empty"), (4099, 20, "dense"), (2, 24, "random"),
])
def test_agg_mask_fast_path_fp32(v2v, B, N, kind):
    rng = np.random.default_rng(B * 131 + N)
    adj = rand_adj(rng, B, N, kind)
    H = rng.normal(size=(B, N, 16)).astype(np.float32)
    im, om, _ = v2v.pack_adjacency(dev(adj))
    out = v2v.aggregate(dev(H), mask=im)
    assert rel_err(out.cpu().numpy(), O.agg_factored(H.astype(np.float64), adj)) <= RTOL_F32
    # transposed orientation == backward w.r.t. H, with an addend (the accumulate form)
    add = rng.normal(size=(B, N, 16)).astype(np.float32)
    outT = v2v.aggregate(dev(H), mask=om, addend=dev(add))
    assert rel_err(outT.cpu().numpy(), O.agg_factored_T(H.astype(np.float64), adj) + add) <= RTOL_F32


def test_agg_addend_may_alias_out(v2v):
    rng = np.random.default_rng(5)
    B, N = 300, 20
    adj = rand_adj(rng, B, N, "dense")
    H = rng.normal(size=(B, N, 16)).astype(np.float32)
    acc = rng.normal(size=(B, N, 16)).astype(np.float32)
    im, _, _ = v2v.pack_adjacency(dev(adj))
    lib = v2v.load_library()
    Hd, accd = dev(H), dev(acc)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.v2v_agg_mask(Hd.data_ptr(), im.data_ptr(), accd.data_ptr(), accd.data_ptr(), B, N, 16, 0, st)
    assert rc == 0, lib.v2v_last_error()
    assert rel_err(accd.cpu().numpy(), O.agg_factored(H.astype(np.float64), adj) + acc) <= RTOL_F32


@pytest.mark.parametrize("B,N,F", [(17, 20, 8), (5, 40, 16), (3, 256, 4), (11, 20, 5), (2, 64, 32)])
def test_agg_mask_generic_path(v2v, B, N, F):
    rng = np.random.default_rng(N * F)
    adj = rand_adj(rng, B, N, "random")
    H = rng.normal(size=(B, N, F)).astype(np.float32)
    im, om, _ = v2v.pack_adjacency(dev(adj))
    assert rel_err(v2v.aggregate(dev(H), mask=im).cpu().numpy(), O.agg_factored(H.astype(np.float64), adj)) <= RTOL_F32
    assert rel_err(v2v.aggregate(dev(H), mask=om).cpu().numpy(), O.agg_factored_T(H.astype(np.float64), adj)) <= RTOL_F32


@pytest.mark.parametrize("B,N", [(64, 20), (1001, 20), (12, 4), (40, 32)])
def test_agg_mask_bf16_storage(v2v, B, N):
    """bf16 storage, fp32 accumulate: the only error is the final round to bf16 (2^-9 relative)."""
    rng = np.random.default_rng(B + N)
    adj = rand_adj(rng, B, N, "dense" if N > 2 else "random")
    H = torch.from_numpy(rng.normal(size=(B, N, 16)).astype(np.float32)).to(torch.bfloat16)
    im, _, _ = v2v.pack_adjacency(dev(adj))
    out = v2v.aggregate(H.cuda(), mask=im).float().cpu().numpy()
    ref = O.agg_factored(H.float().numpy().astype(np.float64), adj)
    assert np.all(np.abs(out - ref) <= 2.0 ** -8 * np.abs(ref) + 1e-6)


def test_agg_weighted_dense_path(v2v):
    rng = np.random.default_rng(2)
    B, N, F = 37, 20, 16
    adj = rng.normal(size=(B, N, N)).astype(np.float32)
    H = rng.normal(size=(B, N, F)).astype(np.float32)
    out = v2v.aggregate(dev(H), adj=dev(adj))
    assert rel_err(out.cpu().numpy(), O.agg_factored(H.astype(np.float64), adj.astype(np.float64))) <= RTOL_F32
    outT = v2v.aggregate(dev(H), adj=dev(adj), transpose=True)
    assert rel_err(outT.cpu().numpy(), O.agg_factored_T(H.astype(np.float64), adj.astype(np.float64))) <= RTOL_F32


@pytest.mark.parametrize("B,N,kind", [
    (300, 33, "random"),      # warp-per-tile walk, W = 2, mask span not 16-byte aligned
    (2000, 64, "dense"),      # warp-per-tile walk, one graph per tile, clear-bit walk
    (777, 48, "sparse2"),     # set-bit walk
    (1500, 100, "dense"),     # CTA-per-graph ring, W = 4, several graphs per CTA
    (2600, 128, "random"),    # more graphs than resident CTAs: both ring stages refill
    (310, 256, "sparse2"), (40, 200, "full"), (9, 97, "empty"), (130, 129, "random"),   # W = 5: unaligned mask rows
])
def test_agg_mask_large_graphs(v2v, B, N, kind):
    """20 < N <= 256: the bit-walk kernels (csrc/agg_kernels.cuh), forward and transposed + addend, fp32 and bf16."""
    rng = np.random.default_rng(B + 7 * N)
    adj = rand_adj(rng, B, N, kind)
    H = rng.normal(size=(B, N, 16)).astype(np.float32)
    im, om, _ = v2v.pack_adjacency(dev(adj))
    ref = O.agg_factored(H.astype(np.float64), adj)
    out = v2v.aggregate(dev(H), mask=im)
    assert rel_err(out.cpu().numpy(), ref) <= RTOL_F32
    add = rng.normal(size=(B, N, 16)).astype(np.float32)
    outT = v2v.aggregate(dev(H), mask=om, addend=dev(add))
    assert rel_err(outT.cpu().numpy(), O.agg_factored_T(H.astype(np.float64), adj) + add) <= RTOL_F32
    Hb = dev(H).to(torch.bfloat16)
    outb = v2v.aggregate(Hb, mask=im)
    refb = O.agg_factored(Hb.float().cpu().numpy().astype(np.float64), adj)
    assert rel_err(outb.float().cpu().numpy(), refb) <= 1e-2                    # bf16 rounding of the stored result


def test_agg_properties_at_full_size(v2v):
    """BASELINE size (B=8192, N=20): linearity and the edge-count checksum, no oracle loop needed."""
    rng = np.random.default_rng(1001)
    B, N = 8192, 20
    adj = rand_adj(rng, B, N, "dense")
    im, om, _ = v2v.pack_adjacency(dev(adj))
    H1 = dev(rng.normal(size=(B, N, 16)).astype(np.float32))
    H2 = dev(rng.normal(size=(B, N, 16)).astype(np.float32))
    a1, a2, a12 = v2v.aggregate(H1, mask=im), v2v.aggregate(H2, mask=im), v2v.aggregate(H1 + 2 * H2, mask=im)
    assert float((a12 - (a1 + 2 * a2)).abs().max()) <= 1e-4
    ones = torch.ones((B, N, 16), device="cuda")
    deg = v2v.aggregate(ones, mask=im)
    assert torch.all(deg == float(N - 2))                       # in-degree N-2 for every node (BS_brain.py:441-445)
    # <Agg(H1), H2> == <H1, AggT(H2)>  (adjointness of forward and backward kernels)
    lhs = float((a1.double() * H2.double()).sum())
    rhs = float((H1.double() * v2v.aggregate(H2, mask=om).double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0)
    # and the oracle on a slice
    sl = slice(4000, 4040)
    assert rel_err(a1[sl].cpu().numpy(), O.agg_factored(H1[sl].cpu().numpy().astype(np.float64), adj[sl])) <= RTOL_F32


def test_agg_empty_batch_and_errors(v2v):
    lib = v2v.load_library()
    assert lib.v2v_agg_mask(None, None, None, None, 0, 20, 16, 0, None) == 0
    H = torch.zeros((2, 4, 16), device="cuda")
    m = torch.zeros((2, 4, 1), dtype=torch.int32, device="cuda")
    assert lib.v2v_agg_mask(H.data_ptr(), m.data_ptr(), None, H.data_ptr(), 2, 4, 16, 0, None) != 0   # out aliases H
    o = torch.zeros_like(H)
    assert lib.v2v_agg_mask(H.data_ptr(), m.data_ptr(), None, o.data_ptr(), 2, 4, 16, 7, None) != 0


# --------------------------------------------------------------------------- dense layers
def _dense_ref(segs, W, b, N, G, act):
    x = np.concatenate([s.astype(np.float64) for s in segs], -1)
    B = x.shape[0] // N
    x = x.reshape(B, N, -1)
    K = x.shape[-1]
    W = W.astype(np.float64)
    out = (x @ W[0, :K] if G == 1 else np.einsum('bnk,nko->bno', x, W[:, :K])) + b.astype(np.float64)[None]
    if act:
        out = np.maximum(out, 0)
    return out.reshape(B * N, -1)


@pytest.mark.parametrize("widths,O_,N,G,B,act", [
    ((9, 4), 16, 20, 1, 1024, 1), ((16, 9, 4, 16), 16, 20, 1, 300, 0), ((9, 16, 16), 80, 20, 1, 257, 1),
    ((80,), 40, 20, 1, 129, 1), ((40,), 20, 20, 1, 64, 1), ((20,), 4, 20, 1, 1000, 0),
    ((9, 4), 16, 4, 4, 77, 1), ((16, 9, 4, 16), 16, 4, 4, 256, 1), ((9, 16, 16), 80, 7, 7, 33, 1), ((80,), 40, 4, 4, 512, 1),
    ((20,), 4, 4, 4, 1, 0), ((9, 4, 16), 16, 1, 1, 50, 1), ((5, 3), 12, 3, 1, 40, 1), ((7,), 6, 5, 5, 21, 0),
])
def test_dense_fwd(v2v, widths, O_, N, G, B, act):
    rng = np.random.default_rng(sum(widths) * O_ + B)
    rows = B * N
    segs = [rng.normal(size=(rows, w)).astype(np.float32) for w in widths]
    K = sum(widths)
    ldw = K + (3 if G == 1 else 0)                        # extra (unused) weight rows, like stage 0's dead W3
    W = rng.normal(size=(G, ldw, O_)).astype(np.float32) * 0.3
    b = rng.normal(size=(G, O_)).astype(np.float32)
    out = v2v.dense_forward([dev(s) for s in segs], dev(W), dev(b), N, G, bool(act))
    assert rel_err(out.cpu().numpy(), _dense_ref(segs, W, b, N, G, act)) <= RTOL_F32


@pytest.mark.parametrize("K,O_,N,G,B,ranges,gate_in,gate_out", [
    (41, 80, 20, 1, 200, ((9, 16), (25, 16)), False, False),
    (45, 16, 20, 1, 333, ((0, 16), (29, 16)), True, False),
    (80, 40, 20, 1, 100, ((0, 80), None), False, True),
    (40, 20, 4, 4, 77, ((0, 40), None), False, True),
    (20, 4, 4, 4, 300, ((0, 20), None), False, True),
    (45, 16, 7, 7, 19, ((0, 16), (29, 16)), True, False),
    (29, 16, 1, 1, 64, ((0, 9), None), True, False),
])
def test_dense_bwd_data(v2v, K, O_, N, G, B, ranges, gate_in, gate_out):
    rng = np.random.default_rng(K * O_ + B)
    lib = v2v.load_library()
    rows = B * N
    dY = rng.normal(size=(rows, O_)).astype(np.float32)
    gin = rng.normal(size=(rows, O_)).astype(np.float32) if gate_in else None
    W = (rng.normal(size=(G, K, O_)) * 0.3).astype(np.float32)
    (k0a, wa), rb = ranges
    k0b, wb = rb if rb else (0, 0)
    gout = rng.normal(size=(rows, wa)).astype(np.float32) if gate_out else None
    dxa = torch.empty((rows, wa), device="cuda")
    dxb = torch.empty((rows, max(wb, 1)), device="cuda")
    dYd, Wd = dev(dY), dev(W)
    gind = dev(gin) if gate_in else None
    goutd = dev(gout) if gate_out else None
    rc = lib.v2v_dense_bwd_data(dYd.data_ptr(), gind.data_ptr() if gate_in else None, Wd.data_ptr(), K, k0a, wa,
                                dxa.data_ptr(), k0b, wb, dxb.data_ptr() if wb else None,
                                goutd.data_ptr() if gate_out else None, B, N, G, O_, None)
    assert rc == 0, lib.v2v_last_error()
    dZ = dY.astype(np.float64) * ((gin > 0) if gate_in else 1.0)
    dZ3 = dZ.reshape(B, N, O_)
    W64 = W.astype(np.float64)
    dX = (dZ3 @ W64[0].T if G == 1 else np.einsum('bno,nko->bnk', dZ3, W64)).reshape(rows, K)
    ra = dX[:, k0a:k0a + wa] * ((gout > 0) if gate_out else 1.0)
    assert rel_err(dxa.cpu().numpy(), ra) <= RTOL_F32
    if wb:
        assert rel_err(dxb.cpu().numpy(), dX[:, k0b:k0b + wb]) <= RTOL_F32


@pytest.mark.parametrize("widths,O_,N,G,B,gate", [
    ((9, 4), 16, 20, 1, 1024, True), ((16, 9, 4, 16), 16, 20, 1, 500, True), ((9, 16, 16), 80, 20, 1, 300, False),
    ((80,), 40, 20, 1, 130, False), ((40,), 20, 20, 1, 64, False), ((20,), 4, 20, 1, 999, False),
    ((16, 9, 4, 16), 16, 4, 4, 256, True), ((9, 16, 16), 80, 7, 7, 33, False), ((5, 3), 6, 3, 1, 40, True),
])
def test_dense_bwd_weight(v2v, widths, O_, N, G, B, gate):
    rng = np.random.default_rng(sum(widths) + O_ * B)
    lib = v2v.load_library()
    rows, K = B * N, sum(widths)
    segs = [rng.normal(size=(rows, w)).astype(np.float32) for w in widths]
    dY = rng.normal(size=(rows, O_)).astype(np.float32)
    g = rng.normal(size=(rows, O_)).astype(np.float32) if gate else None
    ldw = K + 2
    dW = torch.zeros((G, ldw, O_), device="cuda")
    db = torch.zeros((G, O_), device="cuda")
    segd = [dev(s) for s in segs]
    sp = (C.c_void_p * len(segs))(*[t.data_ptr() for t in segd])
    sw = (C.c_int * len(segs))(*widths)
    dYd = dev(dY)
    gd = dev(g) if gate else None
    rc = lib.v2v_dense_bwd_weight(len(segs), sp, sw, dYd.data_ptr(), gd.data_ptr() if gate else None, dW.data_ptr(), ldw,
                                  db.data_ptr(), B, N, G, O_, None)
    assert rc == 0, lib.v2v_last_error()
    x = np.concatenate(segs, -1).astype(np.float64).reshape(B, N, K)
    dZ = (dY.astype(np.float64) * ((g > 0) if gate else 1.0)).reshape(B, N, O_)
    if G == 1:
        rW, rb = np.einsum('bnk,bno->ko', x, dZ)[None], dZ.sum((0, 1))[None]
    else:
        rW, rb = np.einsum('bnk,bno->nko', x, dZ), dZ.sum(0)
    got = dW.cpu().numpy()
    assert rel_err(got[:, :K], rW) <= 2e-5          # fp32 atomics, summation order varies
    assert np.all(got[:, K:] == 0)
    assert rel_err(db.cpu().numpy(), rb) <= 2e-5


# --------------------------------------------------------------------------- loss / target / optimiser
def test_huber_loss_grad(v2v):
    rng = np.random.default_rng(8)
    lib = v2v.load_library()
    B, N, CH = 513, 20, 4
    q = rng.normal(0, 2, (B, N, CH)).astype(np.float32)
    y = rng.normal(0, 2, (B, N, CH)).astype(np.float32)
    dq = torch.empty((B, N, CH), device="cuda")
    hl = torch.zeros(N, device="cuda")
    qd, yd = dev(q), dev(y)
    assert lib.v2v_huber_loss_grad(qd.data_ptr(), yd.data_ptr(), dq.data_ptr(), hl.data_ptr(), B, N, CH, 1.0, None) == 0
    _, per_head = O.brain_loss(q.astype(np.float64), y.astype(np.float64))
    assert rel_err(hl.cpu().numpy(), per_head) <= 1e-5
    ref_dq = np.clip(q.astype(np.float64) - y, -1, 1) / (B * CH)
    assert rel_err(dq.cpu().numpy(), ref_dq) <= 1e-6
    assert abs(v2v.huber_loss(y[:, 0], q[:, 0]) - per_head[0]) <= 1e-5 * per_head[0]


def test_td_target(v2v):
    rng = np.random.default_rng(9)
    lib = v2v.load_library()
    B, N, CH = 300, 20, 4
    p, pn = rng.normal(size=(B, N, CH)).astype(np.float32), rng.normal(size=(B, N, CH)).astype(np.float32)
    a = rng.integers(0, CH, (B, N)).astype(np.int32)
    r = rng.normal(10, 3, B).astype(np.float32)
    y = torch.empty((B, N, CH), device="cuda")
    pd_, pnd, ad, rd = dev(p), dev(pn), torch.from_numpy(a).cuda(), dev(r)
    assert lib.v2v_td_target(pd_.data_ptr(), pnd.data_ptr(), ad.data_ptr(), rd.data_ptr(), 0.5, y.data_ptr(), B, N, CH, None) == 0
    ref = O.td_targets(p, pn, a, r, np.float32(0.5))
    assert np.array_equal(y.cpu().numpy(), ref.astype(np.float32))     # copies + one fp32 fma-free expression


def test_keras_adam_steps(v2v):
    rng = np.random.default_rng(10)
    lib = v2v.load_library()
    n = 10007
    p0 = rng.normal(size=n).astype(np.float32)
    p, m, v = dev(p0), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    rp, rm, rv = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    for t in range(1, 6):
        g = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 2, n)).astype(np.float32)
        gd = dev(g)
        assert lib.v2v_adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, t, 1e-3, 0.5, 0.999, 1e-7, 1.0,
                                 None) == 0
        rp, rm, rv = O.keras_adam_step(rp, g.astype(np.float64), rm, rv, t)
    assert np.abs(p.cpu().numpy() - rp).max() <= 1e-6
    assert rel_err(v.cpu().numpy(), rv) <= 1e-5
