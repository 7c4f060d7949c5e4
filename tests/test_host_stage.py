"""Host staging (csrc/host_stage.cu): strided gather + fp64->fp32 conversion on the worker pool and the host-side
adjacency bit packing, checked against numpy.  No device involved."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib(v2v):
    return v2v.load_library()


def _views(v2v, items):
    L = v2v._lib
    arr = (L.HostView * len(items))(*[L.HostView(*it) for it in items])
    return arr, len(items)


def test_gather_per_slot_fp64_and_kron_adjacency(v2v, lib):
    L = v2v._lib
    rng = np.random.default_rng(0)
    B, N, Dn, F = 37, 5, 9, 16
    slots = [rng.normal(size=(B, Dn)) for _ in range(N)]                       # the reference's D{k}_Node_Input, fp64
    items = [(a.ctypes.data, L.V2V_F64, B, Dn, Dn, 1, k * Dn, N * Dn) for k, a in enumerate(slots)]
    va, n = _views(v2v, items)
    dst = np.full((B, N, Dn), np.nan, np.float32)
    flags = (C.c_int32 * 1)()
    assert lib.v2v_host_gather(va, n, dst.ctypes.data, dst.size, 0, flags) == 0, lib.v2v_last_error()
    assert np.array_equal(dst, np.stack(slots, 1).astype(np.float32))
    # kron(Adj, I_F) (BS_brain.py:492-493) sampled back to (B, N, N) through ONE strided view
    adj = (rng.random((B, N, N)) < 0.6).astype(np.float64)
    A = np.stack([np.kron(a, np.eye(F)) for a in adj])
    va, n = _views(v2v, [(A.ctypes.data, L.V2V_F64, B * N, N, F * N * F, F, 0, N)])
    out = np.empty((B, N, N), np.float32)
    assert lib.v2v_host_gather(va, n, out.ctypes.data, out.size, 1, flags) == 0
    assert np.array_equal(out, adj.astype(np.float32)) and flags[0] == 0
    A[3, 2 * F, 1 * F] = 0.5                                                   # a weighted entry is reported
    assert lib.v2v_host_gather(va, n, out.ctypes.data, out.size, 1, flags) == 0 and (flags[0] & 1)
    # non-zero detection (the reference's all-zero neighbour input is never shipped)
    z = np.zeros((B, F), np.float32)
    va, n = _views(v2v, [(z.ctypes.data, L.V2V_F32, B, F, F, 1, 0, F)])
    o2 = np.empty((B, F), np.float32)
    assert lib.v2v_host_gather(va, n, o2.ctypes.data, o2.size, 2, flags) == 0 and flags[0] == 0
    z[B - 1, F - 1] = 1e-3
    assert lib.v2v_host_gather(va, n, o2.ctypes.data, o2.size, 2, flags) == 0 and (flags[0] & 2)


def test_gather_large_noncontiguous_matches_numpy(v2v, lib):
    L = v2v._lib
    rng = np.random.default_rng(1)
    base = rng.normal(size=(4096, 64)).astype(np.float32)
    win = base[::2, 3:40]                                                       # strided rows, offset columns
    va, n = _views(v2v, [(win.ctypes.data, L.V2V_F32, win.shape[0], win.shape[1], win.strides[0] // 4, 1, 5, 50)])
    dst = np.zeros((win.shape[0], 50), np.float32)
    assert lib.v2v_host_gather(va, n, dst.ctypes.data, dst.size, 0, None) == 0
    assert np.array_equal(dst[:, 5:42], win) and not dst[:, :5].any() and not dst[:, 42:].any()
    # a view that would write past the destination is rejected, nothing is written
    va, n = _views(v2v, [(win.ctypes.data, L.V2V_F32, win.shape[0], win.shape[1], win.strides[0] // 4, 1, 20, 50)])
    assert lib.v2v_host_gather(va, n, dst.ctypes.data, dst.size, 0, None) != 0
    assert b"past the staging tensor" in lib.v2v_last_error()


@pytest.mark.parametrize("N", [1, 4, 7, 20, 31, 32])
@pytest.mark.parametrize("form", ["dense_f32", "dense_f64", "kron_f64"])
def test_host_pack_adjacency_bit_exact(v2v, lib, N, form):
    L = v2v._lib
    rng = np.random.default_rng(N)
    B, F = 50, 3
    adj = (rng.random((B, N, N)) < 0.5).astype(np.float32)
    adj[0] = 0; adj[1] = 1
    if form == "dense_f32":
        src, view = adj, (adj.ctypes.data, L.V2V_F32, B * N, N, N, 1, 0, N)
    elif form == "dense_f64":
        src = adj.astype(np.float64); view = (src.ctypes.data, L.V2V_F64, B * N, N, N, 1, 0, N)
    else:                                                                       # kron(Adj, I_F), BS_brain.py:492-493
        src = np.stack([np.kron(a, np.eye(F)) for a in adj]); view = (src.ctypes.data, L.V2V_F64, B * N, N, F * N * F, F, 0, N)
    va, _ = _views(v2v, [view])
    im = np.zeros((B, N), np.uint32); om = np.zeros((B, N), np.uint32)
    flags = (C.c_int32 * 1)()
    assert lib.v2v_host_pack_adjacency(va, B, N, im.ctypes.data, om.ctypes.data, flags) == 0, lib.v2v_last_error()
    bits = (1 << np.arange(N, dtype=np.uint64))
    want_out = (adj.astype(np.uint64) * bits[None, None, :]).sum(-1).astype(np.uint32)                    # bit m of row n
    want_in = (adj.transpose(0, 2, 1).astype(np.uint64) * bits[None, None, :]).sum(-1).astype(np.uint32)  # bit n of column m
    assert np.array_equal(om, want_out) and np.array_equal(im, want_in) and flags[0] == 0
    k = (N - 1) * F if form == "kron_f64" else N - 1
    src[B - 1, k, k] = 0.25                                                    # last sampled element weighted -> reported
    assert lib.v2v_host_pack_adjacency(va, B, N, im.ctypes.data, om.ctypes.data, flags) == 0 and (flags[0] & 1)
