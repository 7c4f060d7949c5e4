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


def test_select_actions_kernel_follows_the_reference_rule(v2v):
    """v2v_dqn_select_actions against Agent.select_action_while_training (BS_brain.py:308-352) restated in NumPy: linear
    epsilon anneal (:315-324), one explore decision per environment (:330-333), FIRST maximiser on ties (:342-344)."""
    lib, L = v2v.load_library(), v2v._lib
    rng = np.random.default_rng(0)
    E, N, CH = 300, 5, 4
    q = rng.normal(size=(E, N, CH)).astype(np.float32)
    q[::7, :, 2] = q[::7, :, 0] = q[::7].max(axis=2) + 1.0           # ties: columns 0 and 2 share the maximum -> 0 wins
    u = rng.random(E).astype(np.float32)
    rnd = rng.integers(0, CH, (E, N)).astype(np.int32)
    for step, total in ((0, 1000), (400, 1000), (799, 1000), (800, 1000), (5000, 1000)):
        steps = 0.8 * total
        per_step = (1 - 0.01) / steps
        eps = 1 - per_step * step if step < steps else 0.01                              # the reference's branch form
        sched = torch.tensor([step, 1.0, per_step, 0.01], dtype=torch.float32, device="cuda")
        out = torch.empty((E, N), dtype=torch.int32, device="cuda")
        qd, ud, rd = (torch.from_numpy(a).cuda() for a in (q, u, rnd))
        L.check(lib.v2v_dqn_select_actions(qd.data_ptr(), ud.data_ptr(), rd.data_ptr(), sched.data_ptr(), out.data_ptr(), E, N, CH,
                                           L.current_stream()))
        explore = u < np.float32(eps)
        want = np.where(explore[:, None], rnd, q.argmax(axis=2).astype(np.int32))         # np.argmax: first maximiser
        near = np.abs(u - eps) < 1e-6                                                      # float32 rounding of the schedule
        assert np.array_equal(out.cpu().numpy()[~near], want[~near]), step
        assert (want[::7][~explore[::7]] == 0).all()


def test_replay_write_kernel_is_memory_add(v2v):
    """v2v_dqn_replay_write against Memory.add (BS_brain.py:252-256) as the CPU ring implements it: FIFO slots modulo the
    capacity, wrap-around inside one call, device cursor and step counter advanced by the same launch."""
    dqn = _dqn()
    N, cap, Dn, De = 5, 23, 9, 4
    gpu, cpu = dqn.ReplayRing(cap, N, Dn, De, device="cuda"), dqn.ReplayRing(cap, N, Dn, De, device="cpu")
    step = torch.zeros(4, device="cuda")
    g = torch.Generator().manual_seed(0)
    for T in (7, 7, 7, 7, 23, 1):                                  # 4th call wraps (21 + 7 > 23); 5th rewrites every slot
        t = dict(node=torch.randn(T, N, Dn, generator=g), edge=torch.randn(T, N, De, generator=g),
                 in_mask=torch.randint(0, 32, (T, N, 1), generator=g, dtype=torch.int32),
                 out_mask=torch.randint(0, 32, (T, N, 1), generator=g, dtype=torch.int32),
                 action=torch.randint(0, 4, (T, N), generator=g, dtype=torch.int32), reward=torch.randn(T, generator=g),
                 node_=torch.randn(T, N, Dn, generator=g), edge_=torch.randn(T, N, De, generator=g))
        cpu.add_device(**t)
        gpu.add_device(**{k: v.cuda() for k, v in t.items()}, step_dev=step)
        assert gpu.head == cpu.head == int(gpu.head_dev.item()) and gpu.size == cpu.size
        for name in ("node", "edge", "node_", "edge_", "in_mask", "out_mask", "action", "reward"):
            assert torch.equal(getattr(gpu, name).cpu(), getattr(cpu, name)), (T, name)
    assert float(step[0]) == 6.0 and float(step[1:].abs().sum()) == 0.0 and int(gpu._done.item()) == 0
