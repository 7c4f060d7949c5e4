"""CPU tests of the DQN host side: state packing against the restated reference packing on the
UNMODIFIED reference simulator (when /root/reference is present), and the replay ring's bookkeeping."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import v2v_oracle as O


def _dqn():
    from importlib import import_module
    return import_module("globecom2020-resourceallocationgnn_b200.dqn")


@pytest.mark.skipif(not os.path.exists("/root/reference/Environment.py"), reason="reference tree not present")
@pytest.mark.parametrize("n_veh", [4, 20])
def test_pack_state_matches_reference_packing_on_real_simulator(n_veh):
    from oracle import state_packing as SP
    dqn = _dqn()
    env = SP.make_env(n_veh, seed=7)
    for step in range(4):
        node, edge, adj = dqn.pack_state(env, n_veh, env.n_RB)
        ref_state, ref_adj, flat = SP.build_state(env, n_veh, env.n_RB)
        assert np.array_equal(node, ref_state[:, :2 * env.n_RB + 1])
        assert np.array_equal(edge, ref_state[:, 2 * env.n_RB + 1:])
        assert np.array_equal(adj, ref_adj)
        assert flat.shape == (1, n_veh * 13 + n_veh ** 2)                     # BS.num_D2D_Input (:104)
        assert np.array_equal(adj, O.make_adjacency([v.destinations[0] for v in env.vehicles]))
        actions = np.random.randint(0, env.n_RB, (n_veh, 1))
        env.compute_reward_with_channel_selection(actions.copy())
        env.renew_positions(); env.renew_channels_fastfading(); env.Compute_Interference(actions)


def test_replay_ring_fifo_masks_and_sampling_rule():
    dqn = _dqn()
    rng = np.random.default_rng(0)
    N, cap = 5, 12
    ring = dqn.ReplayRing(cap, N, 9, 4, device="cpu")
    store = []
    for t in range(4):                                   # 4 batches of 5 -> wraps around the 12-slot ring
        T = 5
        node, edge, adj, _ = O.synth_batch(T, N, rng)
        node_, edge_, _, _ = O.synth_batch(T, N, rng)
        act = rng.integers(0, 4, (T, N)); rew = rng.normal(size=T)
        ring.add_batch(node, edge, adj, act, rew, node_, edge_)
        store += [(node[i], adj[i], act[i], rew[i], node_[i]) for i in range(T)]
    assert len(ring) == cap and ring.head == 20 % cap
    # slot s holds the most recent transition written there (FIFO eviction of the oldest, BS_brain.py:255-256)
    for s in range(cap):
        j = max(i for i in range(20) if i % cap == s)
        assert np.allclose(ring.node[s].numpy(), store[j][0].astype(np.float32))
        assert np.allclose(ring.node_[s].numpy(), store[j][4].astype(np.float32))
        assert ring.reward[s].item() == pytest.approx(store[j][3], rel=1e-6)
        im, om = O.pack_masks(store[j][1][None])
        assert np.array_equal(ring.in_mask[s].numpy().view(np.uint32), im[0])
        assert np.array_equal(ring.out_mask[s].numpy().view(np.uint32), om[0])
    # sampling rule of Memory.sample (:258-270)
    idx = ring.sample_indices(8, np.random.RandomState(0))
    assert len(set(idx.tolist())) == 8 and idx.max() < cap
    small = dqn.ReplayRing(100, N, 9, 4, device="cpu")
    node, edge, adj, _ = O.synth_batch(3, N, rng)
    small.add_batch(node, edge, adj, rng.integers(0, 4, (3, N)), rng.normal(size=3), node, edge)
    idx = small.sample_indices(16, np.random.RandomState(1))
    assert len(idx) == 16 and idx.max() < 3                 # fewer samples than the batch: drawn with replacement
    g = small.gather(idx)
    assert g["node"].shape == (16, N, 9) and g["action"].dtype == torch.int32
