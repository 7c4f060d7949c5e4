"""The DQN loop against a recording of the reference's OWN loop.

tests/golden/refshim_agent_n4.npz was written by tests/golden/make_refshim_agent_golden.py: the UNMODIFIED `Agent`
(BS_brain.py:280-748) drove the UNMODIFIED simulator for 120 epsilon-greedy transitions and one `replay()`, with
tests/keras_shim standing in for the Keras/TF primitives.  Checked here:

* CPU: the NumPy oracle reproduces what the reference's networks returned, the targets `replay()` built (the TD rule as
  executed, :668-692), the loss and the post-fit weights; `dqn.Agent`'s host loop, run on the real simulator with the
  same seeds and an oracle-backed brain, reproduces the reference's trajectory (states, actions, rewards) -- i.e. the
  same random numbers are consumed in the same order and the state packing is identical;
* GPU: `dqn.Agent` on the CUDA engine, its replay ring loaded with the recorded transitions, reproduces the greedy
  actions and the replay step (History, the four Q statistics, weights after the step).

One behaviour worth knowing: the reference writes the TD value into the array `predict` returned (`t = p[D][b];
t[a] = ...`, :683-690), so its "Orig_Q" statistics (:742-746) are statistics of the TARGETS, equal to Q_mean / Q_max_mean
up to float32 rounding.  The recording shows it, and dqn.Agent reproduces it."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import v2v_oracle as O

PATH = os.path.join(GOLDEN, "refshim_agent_n4.npz")


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def _dqn():
    from importlib import import_module
    return import_module("globecom2020-resourceallocationgnn_b200.dqn")


@pytest.fixture(scope="module")
def rec():
    z = np.load(PATH)
    return {k: z[k] for k in z.files}


def _dims(rec):
    return O.BrainDims(int(rec["N"]), stages=int(rec["S"]), per_slot=bool(rec["per_slot"]))


def test_oracle_reproduces_the_reference_replay(rec):
    d = _dims(rec)
    f64 = lambda a: np.asarray(a, np.float64)
    L, Lt = O.unflatten_params(d, f64(rec["params"])), O.unflatten_params(d, f64(rec["target_params"]))
    idx = rec["replay_index"]
    node, edge, adj = f64(rec["node"][idx]), f64(rec["edge"][idx]), f64(rec["adj"][idx])
    assert rel(O.brain_forward(d, L, node, edge, adj), rec["p"]) <= 1e-4
    assert rel(O.brain_forward(d, Lt, f64(rec["node_"][idx]), f64(rec["edge_"][idx]), adj), rec["p_target"]) <= 1e-4
    # the TD rule as the reference executed it, on the reference's own network outputs (float32 arrays, :683-690)
    y = O.td_targets(f64(rec["p"]), f64(rec["p_target"]), rec["action"][idx], f64(rec["reward"][idx]), float(rec["gamma"]))
    assert np.array_equal(y.astype(np.float32), rec["y"].astype(np.float32))
    loss, per_head, g = O.brain_backward(d, L, node, edge, adj, f64(rec["y"]), q_for_loss=f64(rec["p"]))
    assert abs(loss - float(rec["loss"])) <= 1e-4 * abs(float(rec["loss"]))
    assert rel(per_head, rec["per_head"]) <= 1e-4
    p1, _, _ = O.keras_adam_step(f64(rec["params"]), O.flatten_params(g), 0.0, 0.0, 1)
    assert np.abs(p1 - rec["params_after_fit"]).max() <= 2e-6
    # the statistics replay() returns (:731-746): all four are statistics of the targets (module docstring)
    np.testing.assert_allclose(rec["Q_mean"], rec["y"].mean(axis=(0, 2)), rtol=1e-6)
    np.testing.assert_allclose(rec["Q_max_mean"], rec["y"].max(axis=2).mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(rec["Orig_Q_mean"], rec["Q_mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rec["Orig_Q_max_mean"], rec["Q_max_mean"], rtol=1e-6, atol=1e-7)
    assert np.abs(rec["p"].mean(axis=(0, 2)) - rec["Orig_Q_mean"]).max() > 0.1     # ... and NOT of the original Q values


def test_oracle_reproduces_the_reference_greedy_actions(rec):
    d = _dims(rec)
    L = O.unflatten_params(d, np.asarray(rec["params"], np.float64))
    greedy = ~np.isnan(rec["greedy_q"][:, 0, 0])
    assert greedy.sum() == int(rec["n_greedy"]) >= 20
    q = O.brain_forward(d, L, *(np.asarray(rec[k][greedy], np.float64) for k in ("node", "edge", "adj")))
    assert rel(q, rec["greedy_q"][greedy]) <= 1e-4
    assert np.array_equal(q.argmax(-1), rec["action"][greedy])                       # :342-344


class _OracleBrain:
    """predict_one_step of the recorded weights through the NumPy oracle (CPU stand-in for the CUDA brain)."""

    def __init__(self, rec):
        self.d = _dims(rec)
        self.L = O.unflatten_params(self.d, np.asarray(rec["params"], np.float64))
        self.num_One_Node_Input, self.num_One_Edge_Input = 9, 4

    def predict_one_step(self, x, target=False):
        q = O.brain_forward(self.d, self.L, *(np.asarray(x[k], np.float64) for k in ("Node_Input", "Edge_Input", "Adjacency_Matrix")))
        return [q[:, k] for k in range(self.d.N)]


@pytest.mark.skipif(not os.path.exists("/root/reference/Environment.py"), reason="reference tree not present (GPU box)")
def test_host_loop_reproduces_the_reference_trajectory(rec):
    """dqn.Agent.generate_d2d_transition on the real simulator, same seeds as the recording (RL_Train_main.py:44-47)."""
    from oracle import state_packing as SP
    dqn = _dqn()
    N, CH = int(rec["N"]), int(rec["CH"])
    env = SP.make_env(N, seed=1001)
    agent = dqn.Agent.__new__(dqn.Agent)                  # the constructor builds the CUDA brain; only the host loop is under test
    agent.epsilon, agent.num_step = dqn.MAX_EPSILON, 0
    agent.num_CH, agent.num_D2D, agent.num_Neighbor = CH, N, 1
    agent.env, agent.brain = env, _OracleBrain(rec)
    agent.v2v_weight, agent.v2i_weight = 1, 0.1
    agent.num_Episodes, agent.num_Train_Step, agent.num_transition = 2, 2, 50
    agent.memory = dqn.ReplayRing(1000, N, 9, 4, device="cpu")
    T = rec["node"].shape[0]
    rewards = agent.generate_d2d_transition(T)
    assert agent.epsilon == pytest.approx(float(rec["epsilon_end"]), rel=1e-12)
    np.testing.assert_allclose(rewards, rec["reward"], rtol=1e-12)
    m = agent.memory
    assert np.array_equal(m.action[:T].numpy(), rec["action"])
    for mine, ref in ((m.node, "node"), (m.edge, "edge"), (m.node_, "node_"), (m.edge_, "edge_")):
        assert np.array_equal(mine[:T].numpy(), rec[ref].astype(np.float32))
    im, om = O.pack_masks(rec["adj"])
    assert np.array_equal(m.in_mask[:T].numpy().view(np.uint32), im)
    assert np.array_equal(m.out_mask[:T].numpy().view(np.uint32), om)


@pytest.mark.skipif(not os.path.exists("/root/reference/BS_brain.py"), reason="reference tree not present (GPU box)")
def test_committed_recording_is_what_the_reference_does_now(rec, tmp_path):
    """Fresh interpreter (the fake `keras` never enters this process): re-run the unmodified Agent, compare the files."""
    import subprocess
    import sys
    subprocess.run([sys.executable, os.path.join(GOLDEN, "make_refshim_agent_golden.py"), "/root/reference", str(tmp_path)],
                   check=True, cwd=os.path.dirname(os.path.dirname(GOLDEN)), capture_output=True, timeout=600)
    new = np.load(os.path.join(tmp_path, "refshim_agent_n4.npz"))
    for k, v in rec.items():
        if v.dtype.kind == "f" and k not in ("node", "edge", "node_", "edge_", "adj", "reward", "params", "target_params"):
            np.testing.assert_allclose(new[k], v, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=k)   # float32 network outputs
        else:
            assert np.array_equal(new[k], v), k                                                       # simulator, RNG, weights


class _Cfg:
    Batch_Size, Gamma, v2v_weight, v2i_weight = 64, 0.5, 1, 0.1


# ---------------------------------------------------------------------------------------------------------------------
# the whole training loop: tests/golden/refshim_train_n4.npz = the unmodified Agent.train (BS_brain.py:750-910), 2 episodes
# x 5 train steps x 50 transitions, 10 consecutive Keras-Adam steps and the target synchronisation at env step 500
# (tests/golden/make_refshim_train_golden.py)
TRAIN_PATH = os.path.join(GOLDEN, "refshim_train_n4.npz")


@pytest.fixture(scope="module")
def loop():
    z = np.load(TRAIN_PATH)
    return {k: z[k] for k in z.files}


def test_oracle_follows_the_reference_training_loop(loop):
    """fp64 oracle, its own weights AND its own targets carried from step to step (Adam t = 1..10), against the reference's
    float32 run.  Because each side builds y from its own network output, the untouched entries of the residual are exactly
    zero on both sides (the ill-conditioned Adam steps of tests/test_tf1_golden.py::_check_weights_after_fit never get
    excited) and the two runs stay together to float32 rounding: measured <= 3.3e-6 on y, 7e-7 on the losses, 2.8e-6 on
    every weight after 10 steps."""
    d = _dims(loop)
    f64 = lambda a: np.asarray(a, np.float64)
    P, Tg = f64(loop["params"]), f64(loop["target_params"])
    m, v = np.zeros_like(P), np.zeros_like(P)
    keep = list(loop["params_step_index"])
    ref_head = loop["Train_Loss"].reshape(d.N, -1)                    # (head, episode * steps + step)
    for i in range(len(loop["loss"])):
        idx = loop["replay_index"][i]
        L, Lt = O.unflatten_params(d, P), O.unflatten_params(d, Tg)
        node, edge, adj = f64(loop["node"][idx]), f64(loop["edge"][idx]), f64(loop["adj"][idx])
        p = O.brain_forward(d, L, node, edge, adj)
        p_ = O.brain_forward(d, Lt, f64(loop["node_"][idx]), f64(loop["edge_"][idx]), adj)
        y = O.td_targets(p, p_, loop["action"][idx], f64(loop["reward"][idx]), float(loop["gamma"]))
        assert rel(y, loop["y"][i]) <= 2e-5
        loss, per_head, g = O.brain_backward(d, L, node, edge, adj, y)
        assert abs(loss - loop["loss"][i]) <= 1e-5 * loop["loss"][i]
        assert rel(per_head, ref_head[:, i]) <= 1e-5
        P, m, v = O.keras_adam_step(P, O.flatten_params(g), m, v, i + 1)
        if i in keep:
            assert np.abs(P - loop["params_after_step"][keep.index(i)]).max() <= 1e-5
    assert np.abs(P - loop["params_end"]).max() <= 1e-5
    # the loop's bookkeeping: one synchronisation, at env step 500 after the 10th train step, copies the online weights
    assert loop["sync_at"].tolist() == [[500, 10]]
    assert np.array_equal(loop["params_end"], loop["target_params_end"])
    np.testing.assert_allclose(loop["Orig_Train_Q_mean"], loop["Train_Q_mean"], rtol=1e-5, atol=1e-6)   # in-place overwrite


@pytest.mark.skipif(not os.path.exists("/root/reference/BS_brain.py"), reason="reference tree not present (GPU box)")
def test_committed_training_recording_is_what_the_reference_does_now(loop, tmp_path):
    import subprocess
    import sys
    subprocess.run([sys.executable, os.path.join(GOLDEN, "make_refshim_train_golden.py"), "/root/reference", str(tmp_path)],
                   check=True, cwd=os.path.dirname(os.path.dirname(GOLDEN)), capture_output=True, timeout=900)
    new = np.load(os.path.join(tmp_path, "refshim_train_n4.npz"))
    exact = ("node", "edge", "node_", "edge_", "adj", "reward", "params", "target_params", "action", "replay_index", "sync_at")
    for k, v in loop.items():
        if k in exact or v.dtype.kind != "f":
            assert np.array_equal(new[k], v), k                                    # simulator, RNG draws, indices, weights in
        else:
            np.testing.assert_allclose(new[k], v, rtol=1e-4, atol=1e-5, err_msg=k)  # float32 network arithmetic (BLAS threads)


@pytest.mark.gpu
def test_engine_follows_the_reference_training_loop(loop):
    """dqn.Agent on the CUDA engine replays the reference's 10 train steps: the recorded transitions enter the replay ring
    50 at a time, every replay() draws the recorded batch, the target network is synchronised as Agent.train does."""
    from synthetic_env import SyntheticEnviron
    dqn = _dqn()
    N, CH, F = int(loop["N"]), int(loop["CH"]), int(loop["F"])
    agent = dqn.Agent(N, CH, 1, F, SyntheticEnviron(N, seed=1), _Cfg(), memory_capacity=1000, per_slot=True, seed=0)
    agent.brain.set_flat_params(loop["params"], 0)
    agent.brain.set_flat_params(loop["target_params"], 1)
    ref_head = loop["Train_Loss"].reshape(N, -1)
    q_mean, q_max = loop["Train_Q_mean"].reshape(N, -1), loop["Train_Q_max_mean"].reshape(N, -1)
    for i in range(len(loop["loss"])):
        s = slice(50 * i, 50 * (i + 1))
        agent.memory.add_batch(*(loop[k][s] for k in ("node", "edge", "adj", "action", "reward", "node_", "edge_")))
        agent.num_step = 50 * (i + 1)
        idx = loop["replay_index"][i]
        agent.memory.sample_indices = lambda n, rng=None, idx=idx: idx
        hist, qm, qx, oqm, oqx = agent.replay()
        tol = 1e-4
        assert abs(hist.history["loss"][0] - loop["loss"][i]) <= tol * loop["loss"][i], i
        for k in range(N):
            assert abs(hist.history[f"D{k + 1}_Decide_Output_loss"][0] - ref_head[k, i]) <= tol * ref_head[:, i].max(), (i, k)
        np.testing.assert_allclose(qm, q_mean[:, i], rtol=0, atol=tol * np.abs(q_mean[:, i]).max() + 1e-5)
        np.testing.assert_allclose(qx, q_max[:, i], rtol=0, atol=tol * np.abs(q_max[:, i]).max() + 1e-5)
        if agent.num_step % dqn.UPDATE_TARGET_FREQUENCY == 0:                          # BS_brain.py:846-847
            agent.brain.update_target_model()
    assert agent.brain.iterations == 10
    dev = np.abs(agent.brain.get_flat_params(0) - loop["params_end"])
    print("engine vs reference after 10 steps: median", np.median(dev), "frac > 1e-4", (dev > 1e-4).mean(), "max", dev.max())
    assert dev.max() <= 2e-5                                # measured 2.3e-6 (median 0: most weights bit-equal)
    assert np.array_equal(agent.brain.get_flat_params(1), agent.brain.get_flat_params(0))      # synchronised at step 500


@pytest.mark.gpu
def test_engine_reproduces_the_reference_replay_and_greedy_actions(rec):
    from synthetic_env import SyntheticEnviron
    dqn = _dqn()
    N, CH, F = int(rec["N"]), int(rec["CH"]), int(rec["F"])
    agent = dqn.Agent(N, CH, 1, F, SyntheticEnviron(N, seed=1), _Cfg(), memory_capacity=500, per_slot=True, seed=0)
    agent.brain.set_flat_params(rec["params"], 0)
    agent.brain.set_flat_params(rec["target_params"], 1)
    greedy = np.flatnonzero(~np.isnan(rec["greedy_q"][:, 0, 0]))
    for t in greedy:
        a = agent.greedy_action((rec["node"][t], rec["edge"][t], rec["adj"][t].astype(np.float64)))
        assert np.array_equal(a.reshape(-1), rec["action"][t])
    agent.memory.add_batch(rec["node"], rec["edge"], rec["adj"], rec["action"], rec["reward"], rec["node_"], rec["edge_"])
    idx = rec["replay_index"]
    agent.memory.sample_indices = lambda n, rng=None: idx
    hist, q_mean, q_max, oq_mean, oq_max = agent.replay()
    assert abs(hist.history["loss"][0] - float(rec["loss"])) <= 1e-4 * abs(float(rec["loss"]))
    for k in range(N):
        assert abs(hist.history[f"D{k + 1}_Decide_Output_loss"][0] - rec["per_head"][k]) <= 1e-4 * rec["per_head"].max()
    np.testing.assert_allclose(q_mean, rec["Q_mean"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(q_max, rec["Q_max_mean"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(oq_mean, rec["Orig_Q_mean"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(oq_max, rec["Orig_Q_max_mean"], rtol=1e-4, atol=1e-5)
    # weights after the step: same three-part statement as tests/test_tf1_golden.py::_check_weights_after_fit
    d = _dims(rec)
    f64 = lambda a: np.asarray(a, np.float64)
    _, _, g_ref = O.brain_backward(d, O.unflatten_params(d, f64(rec["params"])), f64(rec["node"][idx]), f64(rec["edge"][idx]),
                                   f64(rec["adj"][idx]), f64(rec["y"]), q_for_loss=f64(rec["p"]))
    g_ref, g_dev = O.flatten_params(g_ref), agent.brain.get_flat_params(2).astype(np.float64)
    scale = np.abs(g_ref).max()
    assert np.abs(g_dev - g_ref).max() <= 1e-5 * scale + 4e-6
    p1 = agent.brain.get_flat_params(0)
    pa, _, _ = O.keras_adam_step(f64(rec["params"]), g_dev, 0.0, 0.0, 1)
    assert np.abs(pa - p1).max() <= 1e-6
    dev, well = np.abs(p1 - rec["params_after_fit"]), np.abs(g_ref) >= 1e-4 * scale
    assert well.mean() > 0.3 and dev[well].max() <= 2e-6 and dev.max() <= 1.1e-3
