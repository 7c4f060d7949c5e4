"""oracle/env_oracle.py (vectorised restatement of the simulator arithmetic) against golden vectors recorded from the
UNMODIFIED reference Environment.py (tests/golden/make_env_golden.py): fp64, 1e-12."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import env_oracle as EO

TOL = 1e-12


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.mark.parametrize("n", [4, 8, 20])
def test_channels_rewards_state_match_reference(n):
    z = load(f"sim_steps_n{n}.npz")
    T = z["pos"].shape[0]
    # the step: channels from the positions AFTER the move, shadows before the update, the reference's own draws
    sh_v, sh_i, v2v_ff, v2i_ff, v2v_abs, v2i_abs = EO.renew_channels(z["pos1"], z["vel"], z["v2v_shadow0"], z["v2i_shadow0"], z["z_v2v"],
                                                                     z["z_v2i"], z["ff_v2v"], z["ff_v2i"])
    assert np.abs(sh_v - z["v2v_shadow1"]).max() <= TOL and np.abs(sh_i - z["v2i_shadow1"]).max() <= TOL
    # channels of step t are the inputs of step t + 1 in the recording (and *_last closes the sequence)
    nxt_v = np.concatenate([z["v2v_ff"][1:], z["v2v_ff_last"]])
    nxt_i = np.concatenate([z["v2i_ff"][1:], z["v2i_ff_last"]])
    assert np.abs(v2v_ff - nxt_v).max() <= 1e-10 and np.abs(v2i_ff - nxt_i).max() <= 1e-10
    assert np.abs(v2v_abs[:-1] - z["v2v_abs"][1:]).max() <= 1e-10 and np.abs(v2i_abs[:-1] - z["v2i_abs"][1:]).max() <= 1e-10
    # rewards on the recorded channels
    v2v_rate, v2i_rate, interf = EO.compute_reward(z["actions"], z["dest"], z["v2v_ff"], z["v2i_ff"], z["v2i_abs"])
    assert np.abs(v2v_rate - z["v2v_rate"]).max() <= 1e-10 * max(1.0, np.abs(z["v2v_rate"]).max())
    assert np.abs(v2i_rate - z["v2i_rate"]).max() <= 1e-10 * max(1.0, np.abs(z["v2i_rate"]).max())
    assert np.abs(interf - z["interference"]).max() <= 1e-12 * max(1e-30, np.abs(z["interference"]).max()) + 1e-300
    # state packing (BS_brain.py:389-407, :441-469)
    node, edge, adj = EO.pack_state(z["dest"], z["v2v_ff"], z["v2i_ff"])
    assert np.abs(node - z["node"]).max() <= 1e-12 and np.abs(edge - z["edge"]).max() <= 1e-12 and np.array_equal(adj, z["adj"])
    assert T >= 4


def test_mobility_matches_reference():
    z = load("sim_mobility.npz")
    cross = EO.crossing(z["pos"], z["dir"], z["vel"])
    # the reference draws lazily, in vehicle order: hand every crossing vehicle its draw
    u = np.ones(cross.shape)
    for t in range(cross.shape[0]):
        draws = z["u"][t][~np.isnan(z["u"][t])]
        assert cross[t].sum() == len(draws), (t, cross[t].sum(), len(draws))
        u[t, cross[t]] = draws
    pos1, dir1 = EO.renew_positions(z["pos"], z["dir"], z["vel"], u)
    assert np.array_equal(dir1, z["dir1"])
    assert np.abs(pos1 - z["pos1"]).max() <= 1e-9
    assert cross.sum() >= 40 and (dir1 != z["dir"]).sum() >= 20                 # the cases do exercise crossings and turns


@pytest.mark.parametrize("n", [4, 8, 20])
def test_destination_candidates_match_reference(n):
    z = load(f"sim_neighbors_n{n}.npz")
    cand = EO.destination_candidates(z["pos"][None])[0]
    assert np.array_equal(cand, z["cand"])
    assert all(z["dest"][i] in cand[i] for i in range(n))                     # the reference's sampled receiver is a candidate
    d = EO.choose_destinations(z["pos"][None], np.full((1, n), 0.999))[0]
    assert np.array_equal(d, cand[:, -1])


def test_steps_mobility_in_recorded_episodes():
    """The ordinary (crossing-free) moves of the recorded episodes."""
    for n in (4, 8, 20):
        z = load(f"sim_steps_n{n}.npz")
        cross = EO.crossing(z["pos"], z["dir"], z["vel"])
        u = np.ones(cross.shape)
        for t in range(cross.shape[0]):
            draws = z["u"][t][~np.isnan(z["u"][t])]
            assert cross[t].sum() == len(draws)
            u[t, cross[t]] = draws
        pos1, dir1 = EO.renew_positions(z["pos"], z["dir"], z["vel"], u)
        assert np.array_equal(dir1, z["dir1"]) and np.abs(pos1 - z["pos1"]).max() <= 1e-9
