"""Batched environment kernels (csrc/env.cu) against golden vectors recorded from the unmodified reference simulator and
against the vectorised oracle; plus full-size properties and a device-resident transition -> brain step."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import env_oracle as EO

pytestmark = pytest.mark.gpu


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def make_env(v2v, T, n):
    env = v2v.BatchedEnviron(T, n_veh=n, n_rb=4, seed=1)
    return env


@pytest.mark.parametrize("n", [4, 8, 20])
def test_channels_rewards_state_against_reference_recording(v2v, n):
    """Every recorded step of the reference is one 'environment' of the batch."""
    z = load(f"sim_steps_n{n}.npz")
    T = z["pos"].shape[0]
    env = make_env(v2v, T, n)
    env.pos.copy_(dev(z["pos1"])); env.vel.copy_(dev(z["vel"]))
    env.v2v_shadow.copy_(dev(z["v2v_shadow0"])); env.v2i_shadow.copy_(dev(z["v2i_shadow0"]))
    env.renew_channels_fastfading(dev(z["z_v2v"]), dev(z["z_v2i"]), dev(z["ff_v2v"]), dev(z["ff_v2i"]))
    nxt_v = np.concatenate([z["v2v_ff"][1:], z["v2v_ff_last"]])
    nxt_i = np.concatenate([z["v2i_ff"][1:], z["v2i_ff_last"]])
    # dB values around 100: fp32 arithmetic, 2e-4 dB absolute
    assert np.abs(env.V2V_channels_with_fastfading.cpu().numpy() - nxt_v).max() <= 2e-4
    assert np.abs(env.V2I_channels_with_fastfading.cpu().numpy() - nxt_i).max() <= 2e-4
    assert np.abs(env.v2v_shadow.cpu().numpy() - z["v2v_shadow1"]).max() <= 1e-5
    assert np.abs(env.v2i_shadow.cpu().numpy() - z["v2i_shadow1"]).max() <= 1e-5
    # rewards and state on the recorded (reference) channels
    env.V2V_channels_with_fastfading.copy_(dev(z["v2v_ff"])); env.V2I_channels_with_fastfading.copy_(dev(z["v2i_ff"]))
    env.V2I_channels_abs.copy_(dev(z["v2i_abs"])); env.dest.copy_(dev(z["dest"], torch.int32))
    v2v_rate, v2i_rate, interf, reward = env.compute_reward_with_channel_selection(dev(z["actions"], torch.int32), 1.0, 0.1)
    assert np.abs(v2v_rate.cpu().numpy() - z["v2v_rate"]).max() <= 1e-4 * max(1.0, np.abs(z["v2v_rate"]).max())
    assert np.abs(v2i_rate.cpu().numpy() - z["v2i_rate"]).max() <= 1e-4 * max(1.0, np.abs(z["v2i_rate"]).max())
    assert np.all(np.abs(interf.cpu().numpy() - z["interference"]) <= 1e-4 * np.abs(z["interference"]) + 1e-30)
    want_r = z["v2v_rate"].sum(1) + 0.1 * z["v2i_rate"].sum(1)
    assert np.abs(reward.cpu().numpy() - want_r).max() <= 1e-4 * np.abs(want_r).max()
    node, edge, im, om, adj = env.pack_state(dense_adj=True)
    assert np.abs(node.cpu().numpy() - z["node"]).max() <= 1e-5 and np.abs(edge.cpu().numpy() - z["edge"]).max() <= 1e-5
    assert np.array_equal(adj.cpu().numpy(), z["adj"])
    rim, rom, binary = v2v.pack_adjacency(dev(z["adj"]))                      # bit-exact against the adjacency packer
    assert binary and torch.equal(im, rim) and torch.equal(om, rom)


def test_mobility_against_reference_recording(v2v):
    z = load("sim_mobility.npz")
    cross = EO.crossing(z["pos"], z["dir"], z["vel"])
    u = np.ones(cross.shape)
    for t in range(cross.shape[0]):
        u[t, cross[t]] = z["u"][t][~np.isnan(z["u"][t])]
    T, n = z["dir"].shape
    env = make_env(v2v, T, n)
    env.pos.copy_(dev(z["pos"])); env.dir.copy_(dev(z["dir"], torch.int32)); env.vel.copy_(dev(z["vel"]))
    env.renew_positions(dev(u))
    assert np.array_equal(env.dir.cpu().numpy(), z["dir1"])
    assert np.abs(env.pos.cpu().numpy() - z["pos1"]).max() <= 2e-4           # fp32 metres on a 1299 m map


@pytest.mark.parametrize("n", [4, 8, 20])
def test_destinations_against_reference_recording(v2v, n):
    z = load(f"sim_neighbors_n{n}.npz")
    cand = z["cand"]
    K = cand.shape[1]
    env = make_env(v2v, K, n)
    env.pos.copy_(dev(np.repeat(z["pos"][None], K, 0)))
    u = (np.arange(K)[:, None] + 0.5) / K * np.ones((K, n))                   # environment k picks candidate k
    env.renew_neighbor(dev(u))
    assert np.array_equal(env.dest.cpu().numpy(), cand.T)


def test_full_size_properties_and_brain_step(v2v):
    """BASELINE-size batch (8192 environments x 20 vehicles): oracle on a slice, invariants on everything, and one
    device-resident transition -> TD target -> train step with no host round trip of the state."""
    E, N = 8192, 20
    env = v2v.BatchedEnviron(E, n_veh=N, n_rb=4, seed=1001)
    env.new_random_game()
    pos0 = env.pos.clone()
    for _ in range(3):
        env.renew_positions(); env.renew_channels_fastfading()
    assert float((env.pos - pos0).abs().max()) <= 3 * 0.15 + 1e-3 + 1299     # moves are <= 0.15 m unless wrapped at the border
    assert bool(((env.pos[..., 0] >= 0) & (env.pos[..., 0] <= 750) & (env.pos[..., 1] >= 0) & (env.pos[..., 1] <= 1299)).all())
    d = env.dest.cpu().numpy()
    assert (d != np.arange(N)[None]).all() and d.min() >= 0 and d.max() < N
    node, edge, im, om, adj = env.pack_state(dense_adj=True)
    assert torch.isfinite(node).all() and torch.isfinite(edge).all()
    assert torch.all(adj.sum(1) == N - 2)                                     # in-degree N-2 (BS_brain.py:441-445)
    sl = slice(100, 132)
    no, eo, ao = EO.pack_state(d[sl], env.V2V_channels_with_fastfading[sl].double().cpu().numpy(),
                               env.V2I_channels_with_fastfading[sl].double().cpu().numpy())
    assert np.abs(node[sl].cpu().numpy() - no).max() <= 1e-5 and np.abs(edge[sl].cpu().numpy() - eo).max() <= 1e-5
    actions = torch.randint(0, 4, (E, N), device="cuda", dtype=torch.int32)
    v2v_rate, v2i_rate, interf, reward = env.compute_reward_with_channel_selection(actions)
    ro = EO.compute_reward(actions[sl].cpu().numpy(), d[sl], env.V2V_channels_with_fastfading[sl].double().cpu().numpy(),
                           env.V2I_channels_with_fastfading[sl].double().cpu().numpy(), env.V2I_channels_abs[sl].double().cpu().numpy())
    assert np.abs(v2v_rate[sl].cpu().numpy() - ro[0]).max() <= 1e-4 * max(1.0, np.abs(ro[0]).max())
    assert torch.isfinite(reward).all() and float(reward.min()) >= 0.0
    # transition -> brain, all on the device
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=2, per_slot=False, max_batch=E, data_parallel=False, seed=3)
    brain.update_target_model()
    q = brain.forward_device(node, edge, in_mask=im)
    env.renew_positions(); env.renew_channels_fastfading()
    node2, edge2, _, _ = env.pack_state()
    q2 = brain.forward_device(node2, edge2, in_mask=im, target=True)
    y = torch.empty_like(q)
    lib = brain._lib
    v2v._lib.check(lib.v2v_td_target(q.data_ptr(), q2.data_ptr(), actions.data_ptr(), reward.data_ptr(), 0.5, y.data_ptr(), E, N, 4,
                                     v2v._lib.current_stream()))
    hl = brain.train_step_device(node, edge, im, om, None, y)
    assert torch.isfinite(hl).all()


def test_batched_dqn_loop_on_device(v2v):
    """E environments x the reference DQN loop, nothing leaves the device but the per-step statistics."""
    class Cfg:                                   # Sim_Config.RL_Config fields the agent reads (Sim_Config.py:8-24)
        Batch_Size, Gamma, v2v_weight, v2i_weight = 256, 0.5, 1.0, 0.1
    E, N = 128, 4
    env = v2v.BatchedEnviron(E, n_veh=N, n_rb=4, seed=5)
    agent = v2v.BatchedAgent(env, Cfg, memory_capacity=8192, seed=7, stages=3, per_slot=True)     # the reference's brain: per-slot weights
    loss, rew = agent.train(num_episodes=2, num_train_steps=6, num_transition=4)
    assert np.isfinite(loss).all() and np.isfinite(rew).all() and (rew > 0).all()
    assert len(agent.memory) == min(8192, 2 * 6 * 4 * E) and agent.num_step == 48
    assert agent.epsilon < 1.0 and agent.brain.iterations == 12
    # one replay step on a FIXED set of ring slots against the oracle (predict x2, TD rule :668-692, Huber, backward):
    # deterministic -- no "the loss went down" assertion on noisy minibatches
    from oracle import v2v_oracle as O
    m, B = agent.memory, Cfg.Batch_Size
    idx = torch.arange(0, 4 * B, 4, device=m.device)[:B]
    g = lambda t: t.index_select(0, idx)
    batch = {k: g(getattr(m, k)).cpu().numpy() for k in ("node", "edge", "node_", "edge_", "in_mask", "action", "reward")}
    d = O.BrainDims(N, stages=3, per_slot=True)
    L = O.unflatten_params(d, agent.brain.get_flat_params(0).astype(np.float64))
    Lt = O.unflatten_params(d, agent.brain.get_flat_params(1).astype(np.float64))
    adj = np.zeros((B, N, N))
    imw = batch["in_mask"].view(np.uint32)[:, :, 0]
    for n in range(N):
        adj[:, n, :] = (imw >> np.uint32(n)) & 1
    f64 = lambda a: a.astype(np.float64)
    p = O.brain_forward(d, L, f64(batch["node"]), f64(batch["edge"]), adj)
    p_ = O.brain_forward(d, Lt, f64(batch["node_"]), f64(batch["edge_"]), adj)
    y = O.td_targets(p, p_, batch["action"], f64(batch["reward"]), Cfg.Gamma)
    want_loss, _, _ = O.brain_backward(d, L, f64(batch["node"]), f64(batch["edge"]), adj, y)
    losses, y_mean, p_mean = agent.replay(indices=idx)
    assert abs(float(losses.sum()) - want_loss) <= 2e-4 * abs(want_loss)
    np.testing.assert_allclose(p_mean.cpu().numpy(), p.mean(axis=(0, 2)), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(y_mean.cpu().numpy(), y.mean(axis=(0, 2)), rtol=1e-4, atol=1e-5)
    assert agent.brain.iterations == 13
    # actions are valid channels, greedy part uses the first maximiser
    node, edge, im, om = env.pack_state()
    agent.epsilon, agent.total_steps, agent.num_step = 0.0, 1, 10
    a = agent.select_actions(node, edge, im)
    q = agent.brain.forward_device(node, edge, in_mask=im)
    assert torch.equal(a, torch.argmax(q, dim=2).to(torch.int32)) and int(a.min()) >= 0 and int(a.max()) < 4


def test_graph_replayed_transitions_equal_eager_transitions(v2v):
    """BatchedAgent.generate_transitions replays ONE captured CUDA graph per environment step; with the same seeds it must
    walk through exactly the states the launch-by-launch path walks through (simulator, replay ring, epsilon, rewards)."""
    class Cfg:
        Batch_Size, Gamma, v2v_weight, v2i_weight = 64, 0.5, 1.0, 0.1
    E, N = 16, 4
    runs = []
    for use_graph in (False, True):
        env = v2v.BatchedEnviron(E, n_veh=N, n_rb=4, seed=11)
        agent = v2v.BatchedAgent(env, Cfg, memory_capacity=256, seed=13, stages=3, per_slot=True, use_graph=use_graph)
        agent.total_steps, agent.num_step = 40, 0
        env.new_random_game()
        r1 = agent.generate_transitions(9)
        losses, _, _ = agent.replay(indices=torch.arange(64, device=env.dev))
        r2 = agent.generate_transitions(12)                          # wraps around the 256-slot ring (21 * 16 = 336 > 256)
        torch.cuda.synchronize()
        assert (agent._graph is not None) == use_graph
        m = agent.memory
        runs.append(dict(r1=r1, r2=r2, losses=losses, eps=agent.epsilon, step=agent.num_step, size=m.size, head=m.head,
                         head_dev=int(m.head_dev.item()), sched=agent._sched.clone(), n_env=env.n_step,
                         ring=[t.clone() for t in (m.node, m.edge, m.node_, m.edge_, m.in_mask, m.out_mask, m.action, m.reward)],
                         env=[t.clone() for t in agent._env_state()]))
    a, b = runs
    assert a["step"] == b["step"] == 21 and a["size"] == b["size"] == 256 and a["head"] == b["head"] == b["head_dev"] == (21 * E) % 256
    assert a["n_env"] == b["n_env"] and a["eps"] == b["eps"] and 0.01 < a["eps"] < 1.0
    assert torch.equal(a["sched"], b["sched"]) and float(a["sched"][0]) == 21.0
    for k in ("r1", "r2", "losses"):
        assert torch.equal(a[k], b[k]), k
    for x, y in zip(a["ring"] + a["env"], b["ring"] + b["env"]):
        assert torch.equal(x, y)
    assert a["r1"].shape == (9, E) and torch.isfinite(a["r2"]).all() and float(a["r1"].min()) >= 0.0
