#!/usr/bin/env python
"""Record the reference's OWN training loop: the unmodified `Agent.train` (BS_brain.py:750-910) -- 2 episodes x 5 train
steps x 50 transitions on the unmodified simulator, batch 64 -- executed in this container on tests/keras_shim.

    python tests/golden/make_refshim_train_golden.py [reference [out]]     # writes tests/golden/refshim_train_n4.npz

500 environment steps, 10 `replay()` calls (10 consecutive Keras-Adam steps: the bias-correction schedule lr_t(t) for
t = 1..10 is exercised), and the target-network synchronisation that fires at env step 500 (UPDATE_TARGET_FREQUENCY,
:275, :846-847).  Same accommodations as make_refshim_agent_golden.py (NumPy < 1.24 behaviour through a proxy bound to
`BS_brain.np`; bound methods of the instances wrapped to copy what flows through them); the loop runs in a scratch
working directory because `train` creates its Windows-style result folder under os.getcwd() (:797-801).

Recorded: initial weights of both networks, all 500 transitions, per train step the sampled batch indices, the targets
fed to `train_dnn` and the History, the arrays `train` returns, the online weights after train steps 1, 5 and 10, both networks' weights at the end."""
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "keras_shim"))
sys.path.insert(0, HERE)
import make_tf1_golden as G                                   # noqa: E402
from make_refshim_agent_golden import split_state, N, F, CH   # noqa: E402

EPISODES, STEPS, BATCH, GAMMA = 2, 5, 64, 0.5
KEEP = (0, 4, 9)          # train steps whose post-step weights are stored (every step would be 1.5 MB)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.abspath(sys.argv[2] if len(sys.argv) > 2 else HERE)
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
    cfg.set_train_value(F, GAMMA, BATCH, 1, 0.1)
    up = [3.5 / 2, 3.5 / 2 + 3.5, 250 + 3.5 / 2, 250 + 3.5 + 3.5 / 2, 500 + 3.5 / 2, 500 + 3.5 + 3.5 / 2]
    down = [250 - 3.5 - 3.5 / 2, 250 - 3.5 / 2, 500 - 3.5 - 3.5 / 2, 500 - 3.5 / 2, 750 - 3.5 - 3.5 / 2, 750 - 3.5 / 2]
    left = [3.5 / 2, 3.5 / 2 + 3.5, 433 + 3.5 / 2, 433 + 3.5 + 3.5 / 2, 866 + 3.5 / 2, 866 + 3.5 + 3.5 / 2]
    right = [433 - 3.5 - 3.5 / 2, 433 - 3.5 / 2, 866 - 3.5 - 3.5 / 2, 866 - 3.5 / 2, 1299 - 3.5 - 3.5 / 2, 1299 - 3.5 / 2]
    env = Environ(down, up, left, right, 750, 1299)
    env.new_random_game(env.n_Veh)
    agent = Agent(N, CH, env.n_Neighbor, F, env, cfg)

    from oracle import v2v_oracle as O
    dims = O.BrainDims(N, 3, 1, F, 1, CH, stages=3, per_slot=True)
    like = O.init_params(dims, np.random.default_rng(0), dtype=np.float32)
    rng = np.random.default_rng(seed + 1)
    for model in (agent.brain.model, agent.brain.target_model):
        layers = G.extract(model, dims, like)
        for l in layers:
            l["b"] += rng.normal(0.0, 0.05, l["b"].shape).astype(np.float32)
        G.inject(model, dims, layers)
    params0 = O.flatten_params(G.extract(agent.brain.model, dims, like))
    target0 = O.flatten_params(G.extract(agent.brain.target_model, dims, like))

    calls = {"train": [], "sample": [], "sync": []}
    brain, memory = agent.brain, agent.memory
    train0, sample0, sync0 = brain.train_dnn, memory.sample, brain.update_target_model

    def train_dnn(data_train, labels, batch_size):
        hist = train0(data_train, labels, batch_size)
        calls["train"].append((np.stack([np.array(labels[f"D{k + 1}_Decide_Output"]) for k in range(N)], 1), hist,
                               O.flatten_params(G.extract(brain.model, dims, like))))
        return hist

    def sample(n):
        batch = sample0(n)
        where = {id(smp[0]): i for i, smp in enumerate(memory.samples)}
        calls["sample"].append(np.array([where[id(b[0])] for b in batch]))
        return batch

    def update_target_model():
        calls["sync"].append((agent.num_step, len(calls["train"])))
        return sync0()

    brain.train_dnn, memory.sample, brain.update_target_model = train_dnn, sample, update_target_model

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as scratch:
        os.chdir(scratch)
        try:
            ret = agent.train(EPISODES, STEPS)
        finally:
            os.chdir(cwd)
    Train_Loss, Reward_Per_Train_Step, Reward_Per_Episode, Q_mean, Q_max_mean, Orig_Q_mean, Orig_Q_max_mean = ret
    T = EPISODES * STEPS * 50
    assert len(memory.samples) == T == agent.num_step and len(calls["train"]) == EPISODES * STEPS
    tr = {k: [] for k in ("node", "edge", "adj", "action", "reward", "node_", "edge_")}
    for s, a, r, s_ in memory.samples:
        node, edge, adj = split_state(s)
        node_, edge_, _ = split_state(s_)
        for k, v in zip(tr, (node, edge, adj, np.asarray(a).reshape(N), r, node_, edge_)):
            tr[k].append(v)
    tr = {k: np.array(v) for k, v in tr.items()}
    assert np.allclose(tr["reward"], Reward_Per_Train_Step.reshape(-1))
    np.savez_compressed(
        os.path.join(out, "refshim_train_n4.npz"), N=N, F=F, CH=CH, S=3, per_slot=1, gamma=GAMMA, batch=BATCH,
        episodes=EPISODES, steps=STEPS, params=params0, target_params=target0,
        node=tr["node"].astype(np.float32), edge=tr["edge"].astype(np.float32), adj=tr["adj"].astype(np.float32),
        action=tr["action"].astype(np.int32), reward=tr["reward"], node_=tr["node_"].astype(np.float32),
        edge_=tr["edge_"].astype(np.float32),
        replay_index=np.stack(calls["sample"]), y=np.stack([c[0] for c in calls["train"]]).astype(np.float32),
        loss=np.array([c[1].history["loss"][0] for c in calls["train"]]),
        params_step_index=np.array(KEEP), params_after_step=np.stack([calls["train"][i][2] for i in KEEP]),
        Train_Loss=Train_Loss, Reward_Per_Episode=Reward_Per_Episode, Train_Q_mean=Q_mean, Train_Q_max_mean=Q_max_mean,
        Orig_Train_Q_mean=Orig_Q_mean, Orig_Train_Q_max_mean=Orig_Q_max_mean,
        sync_at=np.array(calls["sync"]).reshape(-1, 2),
        params_end=O.flatten_params(G.extract(brain.model, dims, like)),
        target_params_end=O.flatten_params(G.extract(brain.target_model, dims, like)), epsilon_end=agent.epsilon)
    print(f"wrote refshim_train_n4.npz: {T} transitions, losses {np.round([c[1].history['loss'][0] for c in calls['train']], 4)}, "
          f"target syncs (env step, train steps done) {calls['sync']}")


if __name__ == "__main__":
    main()
