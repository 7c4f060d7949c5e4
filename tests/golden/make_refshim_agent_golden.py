#!/usr/bin/env python
"""Record the reference's OWN DQN loop: the unmodified `Agent` (BS_brain.py:280-748) driving the unmodified simulator
(Environment.py), executed in this container on tests/keras_shim (see its README for what the stand-in restates).

    python tests/golden/make_refshim_agent_golden.py [reference [out]]     # writes tests/golden/refshim_agent_n4.npz

Seeds and construction follow RL_Train_main.py:44-47, :62-75, :88-92.  Nothing in the reference is edited; the script
only (a) gives the reference module the NumPy < 1.24 behaviour it was written for -- the `np.int` alias (BS_brain.py:352)
and ragged lists becoming object arrays in `np.array(self.samples)` (:262) -- through a proxy bound to `BS_brain.np`, and
(b) wraps the bound methods `brain.predict`, `brain.train_dnn` and `memory.sample` of the *instances* to copy what flows through them.

Recorded: the weights of both networks, 120 epsilon-greedy transitions of `generate_d2d_transition` (:409-553: packed
node/edge state, adjacency, whether the action was greedy, the Q values it saw, action, reward, next state), then one
`replay()` (:555-748): the sampled batch, `predict` of both networks as returned (before the reference overwrites the
taken action's entry in place, :683-690), the targets it fed to `train_dnn`, the History and the four Q statistics it
returns, and the weights after the step.  tests/test_refshim_agent.py holds the NumPy oracle (CPU) and dqn.Agent on the
CUDA engine (GPU) to these vectors."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "keras_shim"))
sys.path.insert(0, HERE)
import make_tf1_golden as G     # noqa: E402  (layer <-> slot mapping, flat parameter layout)

N, F, CH = 4, 16, 4
N_TRANSITIONS, BATCH, GAMMA = 120, 64, 0.5


def split_state(flat):
    """Reference flat state (BS_brain.py:441-469): N x (node 9 | edge 4) then the N x N adjacency."""
    s = np.asarray(flat).reshape(-1)
    per = s[:N * 13].reshape(N, 13)
    return per[:, :9].copy(), per[:, 9:].copy(), s[N * 13:].reshape(N, N).copy()


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else HERE
    sys.path.insert(0, ref)
    import tensorflow as tf
    seed = 1001
    random.seed(seed); np.random.seed(seed); tf.set_random_seed(seed)          # RL_Train_main.py:44-47
    from Environment import Environ
    from Sim_Config import RL_Config
    import BS_brain
    from BS_brain import Agent

    class OldNumpy:
        """numpy as BS_brain.py saw it in 2020: `np.int` exists, ragged nested lists give object arrays."""
        int = int

        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def array(obj, *a, **k):
            try:
                return np.array(obj, *a, **k)
            except ValueError:
                return np.array(obj, *a, dtype=object, **k)

    BS_brain.np = OldNumpy()
    cfg = RL_Config()
    cfg.set_train_value(F, GAMMA, BATCH, 1, 0.1)                                # RL_Train_main.py:57-60 (batch 64 here)
    up = [3.5 / 2, 3.5 / 2 + 3.5, 250 + 3.5 / 2, 250 + 3.5 + 3.5 / 2, 500 + 3.5 / 2, 500 + 3.5 + 3.5 / 2]
    down = [250 - 3.5 - 3.5 / 2, 250 - 3.5 / 2, 500 - 3.5 - 3.5 / 2, 500 - 3.5 / 2, 750 - 3.5 - 3.5 / 2, 750 - 3.5 / 2]
    left = [3.5 / 2, 3.5 / 2 + 3.5, 433 + 3.5 / 2, 433 + 3.5 + 3.5 / 2, 866 + 3.5 / 2, 866 + 3.5 + 3.5 / 2]
    right = [433 - 3.5 - 3.5 / 2, 433 - 3.5 / 2, 866 - 3.5 - 3.5 / 2, 866 - 3.5 / 2, 1299 - 3.5 - 3.5 / 2, 1299 - 3.5 / 2]
    env = Environ(down, up, left, right, 750, 1299)                             # RL_Train_main.py:62-75
    env.new_random_game(env.n_Veh)
    assert env.n_Veh == N and env.n_RB == CH
    agent = Agent(N, CH, env.n_Neighbor, F, env, cfg)
    # epsilon reaches its floor after 0.8 * 2 * 2 * 50 = 160 transitions: both random and greedy actions get recorded
    agent.num_Episodes, agent.num_Train_Step, agent.num_transition = 2, 2, 50

    from oracle import v2v_oracle as O
    dims = O.BrainDims(N, 3, 1, F, 1, CH, stages=3, per_slot=True)
    like = O.init_params(dims, np.random.default_rng(0), dtype=np.float32)
    # glorot weights come from the shim; give the biases a spread so that they matter, then keep what is in the models
    rng = np.random.default_rng(seed)
    for model in (agent.brain.model, agent.brain.target_model):
        layers = G.extract(model, dims, like)
        for l in layers:
            l["b"] += rng.normal(0.0, 0.05, l["b"].shape).astype(np.float32)
        G.inject(model, dims, layers)
    params0 = O.flatten_params(G.extract(agent.brain.model, dims, like))
    target0 = O.flatten_params(G.extract(agent.brain.target_model, dims, like))

    calls = {"predict": [], "train": [], "sample": []}
    brain, memory = agent.brain, agent.memory
    predict0, train0, sample0 = brain.predict, brain.train_dnn, memory.sample

    def predict(data_test, target=False):
        res = predict0(data_test, target=target)
        calls["predict"].append(({k: np.array(v) for k, v in data_test.items()}, bool(target), [np.array(r) for r in res]))
        return res

    def train_dnn(data_train, labels, batch_size):
        calls["train"].append(({k: np.array(v) for k, v in data_train.items()}, {k: np.array(v) for k, v in labels.items()},
                               batch_size))
        return train0(data_train, labels, batch_size)

    def sample(n):
        batch = sample0(n)
        calls["sample"].append(batch)
        return batch

    brain.predict, brain.train_dnn, memory.sample = predict, train_dnn, sample

    rewards = agent.generate_d2d_transition(N_TRANSITIONS)
    assert len(memory.samples) == N_TRANSITIONS and agent.num_step == N_TRANSITIONS
    tr = {k: [] for k in ("node", "edge", "adj", "action", "reward", "node_", "edge_")}
    for s, a, r, s_ in memory.samples:
        node, edge, adj = split_state(s)
        node_, edge_, adj_ = split_state(s_)
        assert np.array_equal(adj, adj_)
        for k, v in zip(tr, (node, edge, adj, np.asarray(a).reshape(N), r, node_, edge_)):
            tr[k].append(v)
    tr = {k: np.array(v) for k, v in tr.items()}
    assert np.allclose(tr["reward"], rewards)
    # the greedy transitions: one predict_one_step each, in order; match them to transitions by their node input
    greedy_q = np.full((N_TRANSITIONS, N, CH), np.nan, np.float32)
    cursor = 0
    for feed, target, res in calls["predict"]:
        assert not target and feed["D1_Node_Input"].shape[0] == 1
        while not np.array_equal(feed["D1_Node_Input"][0], tr["node"][cursor, 0]):
            cursor += 1
        greedy_q[cursor] = np.stack([r[0] for r in res])
        assert np.array_equal(feed["Adjacency_Matrix"][0][::F, ::F], tr["adj"][cursor])
        cursor += 1
    n_greedy = len(calls["predict"])
    calls["predict"].clear()

    hist, q_mean, q_max_mean, orig_q_mean, orig_q_max_mean = agent.replay()
    (feed, tgt0, p), (feed_, tgt1, p_) = calls["predict"]
    assert not tgt0 and tgt1
    x, y, bs = calls["train"][0]
    batch = calls["sample"][0]
    assert bs == BATCH and len(batch) == BATCH
    # which stored transition each batch row is (Memory.sample draws without replacement, :258-270)
    index = np.array([next(i for i, smp in enumerate(memory.samples) if smp[0] is b[0]) for b in batch])
    stack = lambda d, key: np.stack([d[f"D{k + 1}_{key}"] for k in range(N)], 1)
    assert np.array_equal(stack(feed, "Node_Input"), tr["node"][index])
    assert np.array_equal(stack(feed_, "Edge_Input"), tr["edge_"][index])
    assert np.array_equal(feed["Adjacency_Matrix"][:, ::F, ::F], tr["adj"][index])
    assert all(np.array_equal(x[k], feed[k]) for k in feed)
    after = O.flatten_params(G.extract(agent.brain.model, dims, like))
    np.savez_compressed(
        os.path.join(out, "refshim_agent_n4.npz"), N=N, F=F, CH=CH, S=3, per_slot=1, gamma=GAMMA, batch=BATCH,
        params=params0, target_params=target0,
        node=tr["node"], edge=tr["edge"], adj=tr["adj"].astype(np.float32), action=tr["action"].astype(np.int32),
        reward=tr["reward"], node_=tr["node_"], edge_=tr["edge_"], greedy_q=greedy_q, n_greedy=n_greedy,
        replay_index=index, p=np.stack(p, 1), p_target=np.stack(p_, 1), y=stack(y, "Decide_Output"),
        loss=float(hist.history["loss"][0]),
        per_head=np.array([hist.history[f"D{k + 1}_Decide_Output_loss"][0] for k in range(N)]),
        Q_mean=q_mean, Q_max_mean=q_max_mean, Orig_Q_mean=orig_q_mean, Orig_Q_max_mean=orig_q_max_mean,
        params_after_fit=after, epsilon_end=agent.epsilon)
    print(f"wrote refshim_agent_n4.npz: {n_greedy} greedy of {N_TRANSITIONS} transitions, replay loss "
          f"{hist.history['loss'][0]:.6f}, Orig_Q_mean == Q_mean: {np.array_equal(orig_q_mean, q_mean)}")


if __name__ == "__main__":
    main()
