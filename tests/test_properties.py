"""Property tests (hypothesis) of the size-independent facts the path rests on -- CPU only: the oracle against itself in
its two forms, and the host-side adjacency packer of the library (integer work: bit-exact) against the oracle.

SURVEY 8c(3): factored aggregation == the reference's dense Kronecker batch_dot (BS_brain.py:69-76 with A = kron(Adj, I_F),
:492-493) for ANY adjacency, permutation equivariance of the shared-weight network under a relabelling of the nodes, the
TD rule (:668-692) touching exactly the chosen action, mask packing as an exact, invertible encoding of a 0/1 adjacency."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import v2v_oracle as O

FAST = settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@FAST
@given(B=st.integers(1, 4), N=st.integers(1, 9), F=st.sampled_from([1, 3, 16]), seed=st.integers(0, 2**31 - 1),
       weighted=st.booleans())
def test_factored_aggregation_equals_dense_kronecker_form(B, N, F, seed, weighted):
    rng = np.random.default_rng(seed)
    adj = rng.normal(size=(B, N, N)) if weighted else (rng.random((B, N, N)) < 0.6).astype(np.float64)
    H = rng.normal(size=(B, N, F))
    A = np.stack([O.kron_adjacency(a, F) for a in adj])                    # what the reference feeds (:492-493)
    literal = O.agg_layer_call([H[:, k, :] for k in range(N)], A)          # concat -> batch_dot -> slices (:69-76)
    assert np.allclose(np.stack(literal, 1), O.agg_factored(H, adj), rtol=0, atol=1e-12)
    assert np.array_equal(O.adjacency_from_kron(A, F), adj)                # the adapter's strided sampling is exact
    # the backward of the aggregation is the aggregation over the transposed adjacency
    dA = rng.normal(size=(B, N, F))
    assert np.allclose(O.agg_factored_T(dA, adj), O.agg_factored(dA, adj.transpose(0, 2, 1)), rtol=0, atol=1e-12)


@FAST
@given(N=st.integers(2, 8), S=st.integers(1, 3), seed=st.integers(0, 2**31 - 1))
def test_shared_weight_network_is_permutation_equivariant(N, S, seed):
    rng = np.random.default_rng(seed)
    d = O.BrainDims(N, stages=S, per_slot=False)
    L = O.init_params(d, rng, bias_scale=0.1)
    node, edge, adj, _ = O.synth_batch(3, N, rng)
    perm = rng.permutation(N)
    q = O.brain_forward(d, L, node, edge, adj)
    qp = O.brain_forward(d, L, node[:, perm], edge[:, perm], adj[:, perm][:, :, perm])
    assert np.allclose(qp, q[:, perm], rtol=1e-10, atol=1e-10 * np.abs(q).max())


@FAST
@given(B=st.integers(1, 5), N=st.integers(1, 6), seed=st.integers(0, 2**31 - 1), gamma=st.floats(0.0, 1.0))
def test_td_target_changes_exactly_the_chosen_action(B, N, seed, gamma):
    rng = np.random.default_rng(seed)
    CH = 4
    p, pn = rng.normal(size=(B, N, CH)), rng.normal(size=(B, N, CH))
    act = rng.integers(0, CH, (B, N))
    rew = rng.normal(10, 3, B)
    y = O.td_targets(p, pn, act, rew, gamma)
    for b in range(B):
        for k in range(N):
            for c in range(CH):
                want = rew[b] + gamma * pn[b, k].max() if c == act[b, k] else p[b, k, c]
                assert y[b, k, c] == pytest.approx(want, rel=1e-12, abs=1e-12)


@FAST
@given(B=st.integers(1, 5), N=st.integers(1, 32), seed=st.integers(0, 2**31 - 1),
       form=st.sampled_from(["dense_f32", "dense_f64", "kron_f64"]), density=st.floats(0.0, 1.0))
def test_host_packer_is_an_exact_encoding_of_any_binary_adjacency(v2v, B, N, seed, form, density):
    """csrc/host_stage.cu (SSE2 compare + movemask + 32x32 bit transpose) against oracle.pack_masks, and back."""
    L = v2v._lib
    lib = v2v.load_library()
    rng = np.random.default_rng(seed)
    F = 3
    adj = (rng.random((B, N, N)) < density).astype(np.float32)             # self loops and empty rows allowed
    if form == "dense_f32":
        src, view = adj, L.HostView(adj.ctypes.data, L.V2V_F32, B * N, N, N, 1, 0, N)
    elif form == "dense_f64":
        src = adj.astype(np.float64); view = L.HostView(src.ctypes.data, L.V2V_F64, B * N, N, N, 1, 0, N)
    else:
        src = np.stack([np.kron(a, np.eye(F)) for a in adj])
        view = L.HostView(src.ctypes.data, L.V2V_F64, B * N, N, F * N * F, F, 0, N)
    im = np.zeros((B, N), np.uint32); om = np.zeros((B, N), np.uint32)
    flags = (C.c_int32 * 1)()
    assert lib.v2v_host_pack_adjacency((L.HostView * 1)(view), B, N, im.ctypes.data, om.ctypes.data, flags) == 0
    want_in, want_out = O.pack_masks(adj)
    assert np.array_equal(im, want_in[:, :, 0]) and np.array_equal(om, want_out[:, :, 0]) and flags[0] == 0
    # decode: bit n of in_mask[b, m] is adj[b, n, m]; bit m of out_mask[b, n] is adj[b, n, m]
    bits = np.arange(N, dtype=np.uint32)
    dec_in = ((im[:, :, None] >> bits[None, None, :]) & 1).transpose(0, 2, 1)
    dec_out = (om[:, :, None] >> bits[None, None, :]) & 1
    assert np.array_equal(dec_in, adj.astype(np.uint32)) and np.array_equal(dec_out, adj.astype(np.uint32))


def _tf32_trunc(x):
    """The upper 19 bits of an fp32 value (what the TF32 datapath reads), as fp32."""
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)


@FAST
@given(K=st.integers(1, 160), seed=st.integers(0, 2**31 - 1), scale=st.sampled_from([1e-3, 1.0, 1e3]))
def test_three_pass_tf32_products_are_fp32_grade(K, seed, scale):
    """The arithmetic of the tensor-core backward (csrc/fused.cu, tf32_split + the three passes), emulated exactly:
    hi = upper 19 bits, lo = x - hi (exact in fp32), lo read by the tensor core through its upper 19 bits, the three
    products a_lo*b_hi + a_hi*b_lo + a_hi*b_hi of 11-bit significands are exact.  What is lost per product is the
    lo*lo term and the truncation of the two lo operands: at most 3 * 2^-20 of |a||b|, i.e. a dot product carries
    the same order of error as one accumulated in fp32 -- three orders of magnitude below single-pass TF32."""
    rng = np.random.default_rng(seed)
    a = (rng.normal(size=K) * scale).astype(np.float32)
    b = rng.normal(size=K).astype(np.float32)
    ah, bh = _tf32_trunc(a), _tf32_trunc(b)
    al, bl = _tf32_trunc(a - ah), _tf32_trunc(b - bh)
    assert np.array_equal((a - ah).astype(np.float64), a.astype(np.float64) - ah.astype(np.float64))    # lo is exact
    f8 = np.float64
    three = (al.astype(f8) * bh + ah.astype(f8) * bl + ah.astype(f8) * bh).sum()
    exact = (a.astype(f8) * b.astype(f8)).sum()
    bound = np.abs(a.astype(f8) * b.astype(f8)).sum()
    assert abs(three - exact) <= 3 * 2.0 ** -20 * bound
    one = (ah.astype(f8) * bh).sum()                                           # single-pass TF32 for scale
    assert abs(one - exact) <= 2 * 2.0 ** -10 * bound
