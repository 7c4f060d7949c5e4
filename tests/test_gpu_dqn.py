"""GPU tests of the DQN loop on the B200 brain: the device replay step against the host-side oracle
(predict, TD rule of BS_brain.py:668-692, fit) and a short end-to-end training run on a synthetic environment."""
import numpy as np
import pytest
import torch

from oracle import v2v_oracle as O
from synthetic_env import SyntheticEnviron

pytestmark = pytest.mark.gpu


class _Cfg:
    Batch_Size, Gamma, v2v_weight, v2i_weight = 64, 0.5, 1, 0.1


def _dqn():
    from importlib import import_module
    return import_module("globecom2020-resourceallocationgnn_b200.dqn")


@pytest.mark.parametrize("n_veh,per_slot", [(4, True), (20, False)])
def test_replay_step_matches_oracle(n_veh, per_slot):
    dqn = _dqn()
    np.random.seed(3)
    env = SyntheticEnviron(n_veh, seed=1)
    agent = dqn.Agent(n_veh, 4, 1, 16, env, _Cfg(), memory_capacity=500, per_slot=per_slot, seed=11)
    agent.num_Episodes, agent.num_Train_Step = 2, 2
    agent.generate_d2d_transition(80)
    assert len(agent.memory) == 80 and agent.num_step == 80
    # freeze the sampled indices so that the oracle sees the same batch
    idx = agent.memory.sample_indices(_Cfg.Batch_Size, np.random.RandomState(5))
    agent.memory.sample_indices = lambda n, rng=None: idx
    batch = {k: v.cpu().numpy() for k, v in agent.memory.gather(idx).items()}
    d = O.BrainDims(n_veh, stages=3, per_slot=per_slot)
    L = O.unflatten_params(d, agent.brain.get_flat_params(0).astype(np.float64))
    Lt = O.unflatten_params(d, agent.brain.get_flat_params(1).astype(np.float64))
    # adjacency back from the stored in_mask bits
    adj = np.zeros((len(idx), n_veh, n_veh))
    im = batch["in_mask"].view(np.uint32)[:, :, 0]
    for n in range(n_veh):
        adj[:, n, :] = (im >> np.uint32(n)) & 1
    f64 = lambda a: a.astype(np.float64)
    p = O.brain_forward(d, L, f64(batch["node"]), f64(batch["edge"]), adj)
    p_ = O.brain_forward(d, Lt, f64(batch["node_"]), f64(batch["edge_"]), adj)
    y = O.td_targets(p, p_, batch["action"], f64(batch["reward"]), 0.5)
    loss, per_head, g = O.brain_backward(d, L, f64(batch["node"]), f64(batch["edge"]), adj, y)
    hist, q_mean, q_max, oq_mean, oq_max = agent.replay()
    assert abs(hist.history["loss"][0] - loss) <= 2e-4 * abs(loss)
    # the reference overwrites p in place before taking its "Orig_Q" statistics (:683-690, :742-746): they equal the
    # target statistics (tests/test_refshim_agent.py holds the recording that shows it)
    np.testing.assert_allclose(oq_mean, y.mean(axis=(0, 2)), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(oq_max, y.max(axis=2).mean(axis=0), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(q_mean, y.mean(axis=(0, 2)), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(q_max, y.max(axis=2).mean(axis=0), rtol=1e-4, atol=1e-5)
    gr = O.flatten_params(g)
    gg = agent.brain.get_flat_params(2)
    assert np.abs(gg - gr).max() <= 5e-4 * np.abs(gr).max()          # y is built from fp32 Q values on the device


def test_short_training_run_on_synthetic_environment(tmp_path):
    dqn = _dqn()
    np.random.seed(0)
    env = SyntheticEnviron(4, seed=2)
    agent = dqn.Agent(4, 4, 1, 16, env, _Cfg(), memory_capacity=4000, seed=3)
    out = agent.train(num_episodes=6, num_train_steps=5, num_transition=20, save_dir=str(tmp_path), save_model_interval=3)
    Train_Loss, Reward_Per_Train_Step, Reward_Per_Episode = out[0], out[1], out[2]
    assert Train_Loss.shape == (4, 6, 5) and np.isfinite(Train_Loss).all() and np.isfinite(Reward_Per_Episode).all()
    assert agent.num_step == 6 * 5 * 20 and len(agent.memory) == 600
    assert agent.brain.iterations == 30
    assert 0.0 < agent.epsilon < 1.0
    assert (tmp_path / "Q-Network_model_weights-Episode-6-Step-5-Batch-64.npz").exists()
    # target net was synchronised at env step 500 (UPDATE_TARGET_FREQUENCY, BS_brain.py:275, :846) and not since
    assert not np.array_equal(agent.brain.get_flat_params(0), agent.brain.get_flat_params(1))
    rewards = agent.test_run(2, 5)
    assert rewards.shape == (2, 5) and np.isfinite(rewards).all()
