"""Restatement of the reference Agent's state packing (TEST INFRASTRUCTURE ONLY; see
oracle/v2v_oracle.py for the import rules).

Follows ``Agent.get_state`` (BS_brain.py:389-407) and the packing loop of
``generate_d2d_transition`` (BS_brain.py:441-469) literally, against an unmodified
``Environment.Environ`` from /root/reference (which *does* run in this image).
"""
from __future__ import annotations

import numpy as np


def get_state(env, idx, num_d2d):
    """BS_brain.py:389-407."""
    Constant_A = 80
    Constant_B = 60
    dest = env.vehicles[idx[0]].destinations[idx[1]]
    V2V_channel = (env.V2V_channels_with_fastfading[idx[0], dest, :] - Constant_A) / Constant_B
    V2I_channel = (env.V2I_channels_with_fastfading[idx[0], :] - Constant_A) / Constant_B
    V2V_edge = (((np.sum(env.V2V_channels_with_fastfading[:, dest, :], axis=0)
                  - env.V2V_channels_with_fastfading[dest, dest, :])
                 - (num_d2d - 1) * Constant_A) / Constant_B - V2V_channel) / (num_d2d - 2)
    return V2V_channel, V2I_channel, V2V_edge


def build_state(env, num_d2d, num_ch, num_neighbor=1):
    """Returns (D2D_State [N, Dn+De], Adjacency [N,N], flat States (1, N*(Dn+De)+N*N)).

    BS_brain.py:437-469: per node [V2V gain x CH | V2I gain x CH | power | edge x CH].
    """
    N = num_d2d
    power = env.V2V_power_dB_List[env.fixed_v2v_power_index]
    adj = np.ones((N, N)) - np.eye(N)                                     # :441
    for d in range(N):
        for v in range(N):
            if d == env.vehicles[v].destinations[0]:
                adj[d, v] = 0                                            # :444-445
    gi = num_neighbor * num_ch
    state = np.zeros((N, 2 * gi + num_neighbor + gi))
    for d in range(N):
        ch, v2i, edge = get_state(env, [d, 0], N)
        state[d, 0:gi] = ch
        state[d, gi:2 * gi] = v2i
        state[d, 2 * gi:2 * gi + num_neighbor] = power
        state[d, 2 * gi + num_neighbor:] = edge
    flat = np.concatenate((np.reshape(state, [1, -1]), np.reshape(adj, [1, -1])), axis=-1)   # :469
    return state, adj, flat


def make_env(n_veh=4, seed=1001):
    """Environ with the reference's lane geometry (RL_Train_main.py:62-75) and seeds (:44-47)."""
    import random
    import sys
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
    import Environment
    random.seed(seed)
    np.random.seed(seed)
    up = [3.5 / 2, 3.5 / 2 + 3.5, 250 + 3.5 / 2, 250 + 3.5 + 3.5 / 2, 500 + 3.5 / 2, 500 + 3.5 + 3.5 / 2]
    down = [250 - 3.5 - 3.5 / 2, 250 - 3.5 / 2, 500 - 3.5 - 3.5 / 2, 500 - 3.5 / 2, 750 - 3.5 - 3.5 / 2, 750 - 3.5 / 2]
    left = [3.5 / 2, 3.5 / 2 + 3.5, 433 + 3.5 / 2, 433 + 3.5 + 3.5 / 2, 866 + 3.5 / 2, 866 + 3.5 + 3.5 / 2]
    right = [433 - 3.5 - 3.5 / 2, 433 - 3.5 / 2, 866 - 3.5 - 3.5 / 2, 866 - 3.5 / 2, 1299 - 3.5 - 3.5 / 2,
             1299 - 3.5 / 2]
    env = Environment.Environ(down, up, left, right, 750, 1299)
    env.new_random_game(n_veh)
    return env
