"""DQN host loop and replay memory around the B200 brain.

``Agent`` follows the reference ``Agent`` (BS_brain.py:280-910) for the training path: state packing
(:389-407, :441-469), epsilon-greedy action selection (:308-352), environment stepping (:366-376), transition
generation (:409-553), replay with the DQN target rule (:555-748), and the episode loop with target sync and
checkpoints (:750-910).  It drives any object with the reference ``Environ`` interface
(``/root/reference/Environment.py``, unmodified) -- nothing of the simulator is re-implemented here.

What changes, deliberately:
  * ``Memory`` (:245-270, a Python list with O(buffer) ``np.array`` per sample) becomes ``ReplayRing``: packed
    device tensors (features, both adjacency bitmask orientations, actions, rewards), sampled by index gather on
    the device; same sampling rule (without replacement once full enough, with replacement before, :258-270);
  * ``replay`` never round-trips Q values to the host: online/target forwards, the TD target rule (:668-692,
    ``v2v_td_target``) and the fit step all run on device tensors; only the per-head losses and the Q statistics
    (:731-746) come back;
  * ``np.kron`` (:492, :603, :621) is gone: the adjacency travels as N words per graph.
"""
from __future__ import annotations

import datetime
import os

import numpy as np
import torch

from . import _lib
from ._lib import ptr
from .brain import BS, History
from .layers import pack_adjacency

MEMORY_CAPACITY = 1000000          # BS_brain.py:274
UPDATE_TARGET_FREQUENCY = 500      # :275
MAX_EPSILON = 1                    # :276
MIN_EPSILON = 0.01                 # :277


def td_targets_device(lib, p, p_next, action, reward, gamma):
    """DQN target rule (BS_brain.py:668-692) on device tensors: y = p except y[b, k, a[b, k]] = r[b] + gamma * max p_next[b, k].
    The C entry point takes raw pointers, so dtype / layout / shape are checked here."""
    B, N, CH = p.shape
    for name, t, dt, shape in (("p", p, torch.float32, (B, N, CH)), ("p_next", p_next, torch.float32, (B, N, CH)),
                               ("action", action, torch.int32, (B, N)), ("reward", reward, torch.float32, (B,))):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous() or tuple(t.shape) != shape:
            raise ValueError(f"td_targets_device: {name} must be a contiguous CUDA {dt} tensor of shape {shape}, got "
                             f"{t.dtype} {tuple(t.shape)}")
    y = torch.empty_like(p)
    _lib.check(lib.v2v_td_target(ptr(p), ptr(p_next), ptr(action), ptr(reward), float(gamma), ptr(y), B, N, CH,
                                 _lib.current_stream()))
    return y


class ReplayRing:
    """Replay memory stored as ( s, a, r, s_ ) in packed tensors (replaces BS_brain.py:245-270).

    One slot = node/edge features of s and s_, the two adjacency mask orientations of s (the reference re-uses
    the adjacency of s for s_, :583), the action per node and the scalar reward.
    """

    def __init__(self, capacity, num_d2d, node_dim, edge_dim, device=None):
        self.capacity = int(capacity)
        self.N, self.Dn, self.De = int(num_d2d), int(node_dim), int(edge_dim)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        N, W = self.N, (self.N + 31) // 32
        z = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=self.device)
        self.node, self.edge = z(self.capacity, N, self.Dn), z(self.capacity, N, self.De)
        self.node_, self.edge_ = z(self.capacity, N, self.Dn), z(self.capacity, N, self.De)
        self.in_mask, self.out_mask = z(self.capacity, N, W, dt=torch.int32), z(self.capacity, N, W, dt=torch.int32)
        self.action, self.reward = z(self.capacity, N, dt=torch.int32), z(self.capacity)
        self.size = 0          # valid slots
        self.head = 0          # next slot to write (FIFO eviction like samples.pop(0), :255-256)
        # device mirror of `head`: add_device computes its slot indices from it, so that the same launches can be replayed
        # from a CUDA graph (BatchedAgent) without any host value baked in
        self.head_dev = torch.zeros((), dtype=torch.int64, device=self.device)
        self._done = torch.zeros((), dtype=torch.int32, device=self.device)      # scratch of v2v_dqn_replay_write
        self._arange = {}

    def __len__(self):
        return self.size

    def add_batch(self, node, edge, adj, action, reward, node_, edge_):
        """Append T transitions given as host arrays (T leading).  The adjacency is packed on the device."""
        T = len(reward)
        if T == 0:
            return
        dev = self.device
        to = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a)).to(dev, dtype=dt, non_blocking=True)
        adj_d = to(adj)
        if dev.type == "cuda":
            in_m, out_m, binary = pack_adjacency(adj_d)
            if not binary:
                raise ValueError("ReplayRing stores 0/1 adjacency only")
        else:                       # CPU tensors: host-side bookkeeping tests only, same bit layout
            nz = adj_d != 0
            bits = (1 << torch.arange(self.N, dtype=torch.int64))
            in_m = (nz.permute(0, 2, 1).to(torch.int64) * bits).sum(-1).to(torch.int32).unsqueeze(-1)
            out_m = (nz.to(torch.int64) * bits).sum(-1).to(torch.int32).unsqueeze(-1)
        idx = (self.head + torch.arange(T)) % self.capacity
        idx_d = idx.to(dev)
        for dst, src in ((self.node, to(node)), (self.edge, to(edge)), (self.node_, to(node_)), (self.edge_, to(edge_)),
                         (self.in_mask, in_m), (self.out_mask, out_m), (self.action, to(action, torch.int32)),
                         (self.reward, to(reward))):
            dst.index_copy_(0, idx_d, src)
        self._advance(T)
        self.head_dev.fill_(self.head)

    def _advance(self, T):
        """Host-side bookkeeping of T appended transitions (the device cursor is advanced by add_device itself)."""
        self.head = int((self.head + T) % self.capacity)
        self.size = min(self.capacity, self.size + T)

    def add_device(self, node, edge, in_mask, out_mask, action, reward, node_, edge_, step_dev=None):
        """Append T transitions that already live on the device (the batched environment's output): no host copy.
        ``step_dev`` (fp32 device tensor, optional): element 0 is incremented by the same launch (the agent's step counter)."""
        T = int(reward.shape[0])
        if T == 0:
            return
        if T > self.capacity:
            raise ValueError("more transitions than the ring holds")
        for name, t, ref in (("node", node, self.node), ("edge", edge, self.edge), ("node_", node_, self.node_),
                             ("edge_", edge_, self.edge_), ("in_mask", in_mask, self.in_mask), ("out_mask", out_mask, self.out_mask),
                             ("reward", reward, self.reward)):
            if t.device != ref.device or t.dtype != ref.dtype or t.numel() != T * ref[0].numel():
                raise ValueError(f"ReplayRing.add_device: {name} must be {ref.dtype} on {ref.device} with {T} x "
                                 f"{tuple(ref.shape[1:])} elements, got {t.dtype} {tuple(t.shape)} on {t.device}")
        action = action.to(torch.int32)
        if self.device.type == "cuda":
            # one launch writes all eight tensors at the device-side cursor and advances it (csrc/dqn.cu)
            srcs = [t.contiguous() for t in (node, edge, node_, edge_, in_mask, out_mask, action, reward)]
            lib = _lib.load()
            _lib.check(lib.v2v_dqn_replay_write(
                ptr(self.node), ptr(self.edge), ptr(self.node_), ptr(self.edge_), ptr(self.in_mask), ptr(self.out_mask),
                ptr(self.action), ptr(self.reward), ptr(srcs[0]), ptr(srcs[1]), ptr(srcs[2]), ptr(srcs[3]), ptr(srcs[4]),
                ptr(srcs[5]), ptr(srcs[6]), ptr(srcs[7]), ptr(self.head_dev), ptr(step_dev), ptr(self._done), T, self.capacity,
                self.N, self.Dn, self.De, _lib.current_stream()))
        else:
            if T not in self._arange:
                self._arange[T] = torch.arange(T, device=self.device)
            idx = (self.head_dev + self._arange[T]) % self.capacity
            for dst, src in ((self.node, node), (self.edge, edge), (self.node_, node_), (self.edge_, edge_), (self.in_mask, in_mask),
                             (self.out_mask, out_mask), (self.action, action), (self.reward, reward)):
                dst.index_copy_(0, idx, src.reshape((T,) + tuple(dst.shape[1:])))
            self.head_dev.add_(T).remainder_(self.capacity)
            if step_dev is not None:
                step_dev[0] += 1.0
        self._advance(T)

    def sample_indices(self, n, rng=np.random):
        """Memory.sample (:258-270): without replacement when enough samples exist, else with replacement."""
        if self.size >= n:
            return rng.choice(self.size, n, replace=False)
        return rng.randint(0, self.size, size=n)

    def gather(self, indices):
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64)).to(self.device)
        g = lambda t: t.index_select(0, idx)
        return {"node": g(self.node), "edge": g(self.edge), "node_": g(self.node_), "edge_": g(self.edge_),
                "in_mask": g(self.in_mask), "out_mask": g(self.out_mask), "action": g(self.action), "reward": g(self.reward)}


def get_state(env, idx, num_d2d):
    """Agent.get_state (BS_brain.py:389-407): normalised V2V gain, V2I gain and the edge feature of one link."""
    A, Bc = 80, 60
    dest = env.vehicles[idx[0]].destinations[idx[1]]
    v2v = (env.V2V_channels_with_fastfading[idx[0], dest, :] - A) / Bc
    v2i = (env.V2I_channels_with_fastfading[idx[0], :] - A) / Bc
    edge = (((np.sum(env.V2V_channels_with_fastfading[:, dest, :], axis=0)
              - env.V2V_channels_with_fastfading[dest, dest, :]) - (num_d2d - 1) * A) / Bc - v2v) / (num_d2d - 2)
    return v2v, v2i, edge


def pack_state(env, num_d2d, num_ch):
    """Per-node features [V2V gain x CH | V2I gain x CH | power] (node) and [edge x CH] (edge) plus the adjacency
    of BS_brain.py:441-445 (Adj = 1 - I, Adj[n, m] = 0 where n is m's receiver)."""
    N = num_d2d
    power = env.V2V_power_dB_List[env.fixed_v2v_power_index]
    node = np.empty((N, 2 * num_ch + 1), np.float64)
    edge = np.empty((N, num_ch), np.float64)
    adj = np.ones((N, N)) - np.eye(N)
    for d in range(N):
        v2v, v2i, e = get_state(env, [d, 0], N)
        node[d, :num_ch], node[d, num_ch:2 * num_ch], node[d, 2 * num_ch] = v2v, v2i, power
        edge[d] = e
        adj[env.vehicles[d].destinations[0], d] = 0
    return node, edge, adj


class Agent:
    """Define the BS Agent class -- training path of the reference ``Agent`` (BS_brain.py:280-910)."""

    def __init__(self, num_d2d, num_ch, num_neighbor, num_d2d_feedback, environment, curr_rl_config,
                 memory_capacity=MEMORY_CAPACITY, **brain_kwargs):
        if num_neighbor != 1:
            raise ValueError("the reference runs with one neighbour per V2V pair (Environment.py:207)")
        self.epsilon = MAX_EPSILON
        self.num_step = 0
        self.num_CH, self.num_D2D, self.num_Neighbor, self.num_Feedback = num_ch, num_d2d, num_neighbor, num_d2d_feedback
        self.input_Node_Info, self.input_Edge_Info = 3, 1                               # :294-295
        self.env = environment
        brain_kwargs.setdefault("max_batch", max(curr_rl_config.Batch_Size, 64))
        self.brain = BS(num_d2d, self.input_Node_Info, self.input_Edge_Info, num_d2d_feedback, num_neighbor, num_ch,
                        **brain_kwargs)
        self.num_States = self.brain.num_D2D_Input
        self.num_Actions = num_ch * num_neighbor
        self.batch_size, self.gamma = curr_rl_config.Batch_Size, curr_rl_config.Gamma
        self.v2v_weight, self.v2i_weight = curr_rl_config.v2v_weight, curr_rl_config.v2i_weight
        self.memory = ReplayRing(memory_capacity, num_d2d, self.brain.num_One_Node_Input, self.brain.num_One_Edge_Input)
        self.num_Episodes, self.num_Train_Step, self.num_transition = 1, 1, 50
        self._lib = _lib.load()

    # ------------------------------------------------------------------ acting
    def _update_epsilon(self):
        """Linear anneal over 80% of all environment steps (:315-324)."""
        steps = self.num_Episodes * 0.8 * self.num_Train_Step * self.num_transition
        per_step = (MAX_EPSILON - MIN_EPSILON) / steps
        self.epsilon = MAX_EPSILON - per_step * self.num_step if self.num_step < steps else MIN_EPSILON

    def select_action_while_training(self, state):
        """state = (node [N,Dn], edge [N,De], adj [N,N]).  Returns an int action matrix (N, num_neighbor) (:308-352)."""
        self._update_epsilon()
        if np.random.random() < self.epsilon:
            acts = np.zeros((self.num_D2D, self.num_Neighbor))
            for d in range(self.num_D2D):
                acts[d, :] = np.random.choice(range(self.num_CH), self.num_Neighbor)
            return acts.astype(int)
        return self.greedy_action(state)

    def greedy_action(self, state):
        node, edge, adj = state
        q = self.brain.predict_one_step({"Node_Input": node[None], "Edge_Input": edge[None], "Adjacency_Matrix": adj[None]})
        # np.where(q == max) keeps the FIRST maximiser (:342-344)
        return np.array([[int(np.argmax(q[d][0]))] for d in range(self.num_D2D)], dtype=int)

    def act(self, actions):
        """Agent.act (:366-376)."""
        self.num_step += 1
        v2v_rate, v2i_rate, interference = self.env.compute_reward_with_channel_selection(actions)
        self.env.renew_positions()
        self.env.renew_channels_fastfading()
        self.env.Compute_Interference(actions)
        return v2v_rate, v2i_rate, interference

    def generate_d2d_transition(self, num_transitions):
        """Epsilon-greedy roll-out of ``num_transitions`` steps into the replay ring (:409-553)."""
        N = self.num_D2D
        rewards = np.zeros(num_transitions)
        rec = {k: [] for k in ("node", "edge", "adj", "action", "reward", "node_", "edge_")}
        for t in range(num_transitions):
            node, edge, adj = pack_state(self.env, N, self.num_CH)
            action = self.select_action_while_training((node, edge, adj))
            v2v_rate, v2i_rate, _ = self.act(action.copy())
            reward = self.v2v_weight * np.sum(np.sum(v2v_rate, axis=1)) + self.v2i_weight * np.sum(v2i_rate)   # :513-517
            rewards[t] = reward
            node_, edge_, _ = pack_state(self.env, N, self.num_CH)          # next state re-uses the adjacency (:545, :583)
            for k, v in (("node", node), ("edge", edge), ("adj", adj), ("action", action.reshape(-1)), ("reward", reward),
                         ("node_", node_), ("edge_", edge_)):
                rec[k].append(v)
        self.memory.add_batch(np.stack(rec["node"]), np.stack(rec["edge"]), np.stack(rec["adj"]), np.stack(rec["action"]),
                              np.asarray(rec["reward"]), np.stack(rec["node_"]), np.stack(rec["edge_"]))
        return rewards

    # ------------------------------------------------------------------ learning
    def replay(self):
        """One replay step on the device (:555-748).  Returns (History, Q_mean, Q_max_mean, Orig_Q_mean, Orig_Q_max_mean)."""
        B, N, CH = self.batch_size, self.num_D2D, self.num_CH
        batch = self.memory.gather(self.memory.sample_indices(B))
        brain = self.brain
        p = brain.forward_device(batch["node"], batch["edge"], in_mask=batch["in_mask"])                      # :664
        p_ = brain.forward_device(batch["node_"], batch["edge_"], in_mask=batch["in_mask"], target=True)       # :665
        y = td_targets_device(self._lib, p, p_, batch["action"], batch["reward"], self.gamma)                  # :668-692
        losses = brain.train_step_device(batch["node"], batch["edge"], batch["in_mask"], batch["out_mask"], None, y)  # :728
        # The reference writes the TD value INTO the array predict returned (`t = p[D][b]; t[a] = ...`, :683-690), so by the
        # time it computes its "Orig_Q" statistics (:742-746) p equals the targets: all four statistics are statistics of
        # y.  Reproduced as executed (recording of the unmodified Agent: tests/test_refshim_agent.py).
        stats = torch.stack([y.mean(dim=(0, 2)), y.max(dim=2).values.mean(dim=0),
                             y.mean(dim=(0, 2)), y.max(dim=2).values.mean(dim=0), losses]).cpu().numpy()          # :731-746
        hist = History()
        hist.history = {"loss": [float(stats[4].sum())]}
        for k in range(N):
            hist.history[f"D{k + 1}_Decide_Output_loss"] = [float(stats[4][k])]
        return hist, stats[0], stats[1], stats[2], stats[3]

    def train(self, num_episodes, num_train_steps, num_transition=50, save_dir=None, save_model_interval=5, verbose=False):
        """Agent.train (:750-910).  Checkpoints (``.npz``) are written only when ``save_dir`` is given."""
        self.num_Episodes, self.num_Train_Step, self.num_transition = num_episodes, num_train_steps, num_transition
        N = self.num_D2D
        Train_Loss = np.ones((N, num_episodes, num_train_steps))
        Train_Q_mean = np.zeros((N, num_episodes, num_train_steps))
        Train_Q_max_mean = np.zeros((N, num_episodes, num_train_steps))
        Orig_Train_Q_mean = np.zeros((N, num_episodes, num_train_steps))
        Orig_Train_Q_max_mean = np.zeros((N, num_episodes, num_train_steps))
        self.num_step = 0
        Reward_Per_Episode = np.zeros(num_episodes)
        Reward_Per_Train_Step = np.zeros((num_episodes, num_train_steps, num_transition))
        for ep in range(num_episodes):
            self.env.new_random_game(N)                                                     # :810
            if verbose and (ep + 1) % 200 == 0:
                print(datetime.datetime.now().strftime('%Y/%m/%d %H:%M:%S'), 'episode', ep + 1, '/', num_episodes)
            for it in range(num_train_steps):
                Reward_Per_Train_Step[ep, it, :] = self.generate_d2d_transition(num_transition)  # :827
                hist, q_mean, q_max, oq_mean, oq_max = self.replay()                              # :832
                for d in range(N):
                    Train_Loss[d, ep, it] = hist.history[f"D{d + 1}_Decide_Output_loss"][0]
                Train_Q_mean[:, ep, it], Train_Q_max_mean[:, ep, it] = q_mean, q_max
                Orig_Train_Q_mean[:, ep, it], Orig_Train_Q_max_mean[:, ep, it] = oq_mean, oq_max
                if self.num_step % UPDATE_TARGET_FREQUENCY == 0:                              # :846-847
                    self.brain.update_target_model()
            Reward_Per_Episode[ep] = np.sum(Reward_Per_Train_Step[ep])
            if save_dir is not None and (ep + 1) % save_model_interval == 0:                  # :853-870
                os.makedirs(save_dir, exist_ok=True)
                tag = f"-Episode-{ep + 1}-Step-{num_train_steps}-Batch-{self.batch_size}"
                self.brain.model.save_weights(os.path.join(save_dir, "Q-Network_model_weights" + tag))
                self.brain.target_model.save_weights(os.path.join(save_dir, "Target-Network_model_weights" + tag))
        return (Train_Loss, Reward_Per_Train_Step, Reward_Per_Episode, Train_Q_mean, Train_Q_max_mean, Orig_Train_Q_mean,
                Orig_Train_Q_max_mean)

    def test_run(self, num_episodes, num_test_steps):
        """Greedy roll-outs with the trained Q network (the core of Agent.test_run, :986-1160, without the
        brute-force optimum and the random baseline).  Returns the per-step reward array."""
        N = self.num_D2D
        out = np.zeros((num_episodes, num_test_steps))
        for ep in range(num_episodes):
            self.env.new_random_game(N)
            for t in range(num_test_steps):
                state = pack_state(self.env, N, self.num_CH)
                action = self.greedy_action(state)                                             # :1108-1115
                v2v_rate, v2i_rate, _ = self.act(action.copy())
                out[ep, t] = self.v2v_weight * np.sum(np.sum(v2v_rate, axis=1)) + self.v2i_weight * np.sum(v2i_rate)
        return out


class BatchedAgent:
    """The reference DQN loop (BS_brain.py:409-553 transitions, :555-748 replay, :750-910 episodes) over E environments
    at once, entirely on the device: ``BatchedEnviron`` (csrc/env.cu) produces states and rewards, the brain acts on
    them with ``forward_device``, transitions go into the ``ReplayRing`` by index copy, and ``replay`` is the same
    device-side step as ``Agent.replay``.  One call of ``generate_transitions(T)`` adds E*T transitions.

    Action selection follows :308-352 per environment: with probability epsilon (linear anneal, :315-324) all N links of
    that environment draw a uniform random channel, otherwise each link takes the first maximiser of its Q row.
    """

    def __init__(self, environment, curr_rl_config, num_d2d_feedback=16, memory_capacity=1 << 18, seed=None, use_graph=True,
                 **brain_kwargs):
        self.env = environment
        self.num_D2D, self.num_CH, self.num_Neighbor = environment.n_Veh, environment.n_RB, 1
        self.E = environment.E
        self.num_step = 0
        brain_kwargs.setdefault("max_batch", max(curr_rl_config.Batch_Size, self.E))
        if seed is not None:
            brain_kwargs.setdefault("seed", int(seed))      # the same seed fixes the glorot draw (reproducible runs)
        self.brain = BS(self.num_D2D, 3, 1, num_d2d_feedback, 1, self.num_CH, **brain_kwargs)
        self.batch_size, self.gamma = curr_rl_config.Batch_Size, curr_rl_config.Gamma
        self.v2v_weight, self.v2i_weight = curr_rl_config.v2v_weight, curr_rl_config.v2i_weight
        self.memory = ReplayRing(memory_capacity, self.num_D2D, self.brain.num_One_Node_Input, self.brain.num_One_Edge_Input,
                                 device=environment.dev)
        self.gen = torch.Generator(device=environment.dev)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        self._lib = _lib.load()
        # The epsilon schedule lives on the device ([step, anneal steps, decrement per step, forced value or -1]): a
        # transition reads nothing from the host, so the whole transition can be replayed from a CUDA graph.
        self._sched = torch.zeros(4, dtype=torch.float32, device=environment.dev)
        self._forced_eps = None
        self.total_steps = 1
        self.use_graph = bool(use_graph)
        self._graph = None               # (CUDAGraph, static reward tensor, library kernels per replay) of one transition
        self._graph_key = None
        self.replayed_kernel_launches = 0        # launches of this library's kernels executed by graph replays (the library's
                                                 # own counter, v2v_launch_count, only sees the launch that was captured)

    # ------------------------------------------------------------------ epsilon schedule (:315-324), host view + device copy
    @property
    def total_steps(self):
        return self._total_steps

    @total_steps.setter
    def total_steps(self, v):
        self._total_steps = v
        self._push_schedule()

    @property
    def num_step(self):
        return self._num_step

    @num_step.setter
    def num_step(self, v):
        self._num_step = int(v)
        if hasattr(self, "_sched"):
            self._push_schedule()

    @property
    def epsilon(self):
        if self._forced_eps is not None:
            return self._forced_eps
        steps = 0.8 * self._total_steps
        per_step = (MAX_EPSILON - MIN_EPSILON) / max(steps, 1)
        return MAX_EPSILON - per_step * self._num_step if self._num_step < steps else MIN_EPSILON

    @epsilon.setter
    def epsilon(self, v):                # pins epsilon (tests, greedy evaluation); None returns to the schedule
        self._forced_eps = None if v is None else float(v)
        self._push_schedule()

    def _push_schedule(self):
        """Device copy of the schedule: {step, base, decrement per step, floor}, epsilon = max(floor, base - decrement * step)
        -- the linear anneal of :315-324 (base - decrement * step reaches the floor exactly when the anneal ends)."""
        if not hasattr(self, "_sched") or not hasattr(self, "_total_steps"):
            return
        steps = 0.8 * self._total_steps
        per_step = (MAX_EPSILON - MIN_EPSILON) / max(steps, 1)
        vals = [float(self._num_step), MAX_EPSILON, per_step, MIN_EPSILON]
        if self._forced_eps is not None:
            vals = [float(self._num_step), self._forced_eps, 0.0, self._forced_eps]
        self._sched.copy_(torch.tensor(vals, dtype=torch.float32))

    def select_actions(self, node, edge, in_mask):
        """[E, N] int32 channel per link (epsilon-greedy per environment, first maximiser on ties: :342-344); the epsilon
        test, the arg-max and the selection are one launch (v2v_dqn_select_actions, csrc/dqn.cu)."""
        E, N, dev = self.E, self.num_D2D, self.env.dev
        q = self.brain.forward_device(node, edge, in_mask=in_mask)
        u = torch.rand((E,), generator=self.gen, device=dev)
        rnd = torch.randint(0, self.num_CH, (E, N), generator=self.gen, device=dev, dtype=torch.int32)
        actions = torch.empty((E, N), dtype=torch.int32, device=dev)
        _lib.check(self._lib.v2v_dqn_select_actions(ptr(q), ptr(u), ptr(rnd), ptr(self._sched), ptr(actions), E, N, self.num_CH,
                                                    _lib.current_stream()))
        return actions

    def _transition(self):
        """One environment step of every environment into the replay ring (:409-553); everything it reads -- simulator
        state, weights, epsilon schedule, ring cursor, generator states -- lives on the device.  Returns the reward [E]."""
        node, edge, im, om = self.env.pack_state()
        actions = self.select_actions(node, edge, im)
        _, _, _, reward = self.env.act(actions, self.v2v_weight, self.v2i_weight)           # :366-376, :513-519
        node_, edge_, _, _ = self.env.pack_state()                                           # the adjacency of s is re-used (:545, :583)
        self.memory.add_device(node, edge, im, om, actions, reward, node_, edge_, step_dev=self._sched)   # also: step += 1
        return reward

    def _capture(self):
        """One transition as a CUDA graph: ~35 small launches (simulator kernels, the brain's forward, action selection,
        ring writes) become one graph launch -- the loop is launch-bound, not compute-bound."""
        key = (self.E, self.num_D2D, self.brain._handle, self.memory.capacity)
        if self._graph is not None and self._graph_key == key:
            return self._graph
        # 1. one eager transition off to the side: lazy initialisation (program upload, kernel attributes) must not happen
        #    inside a capture.  Everything it touched is put back, so that the graph path and the eager path walk through
        #    the same states (identical trajectories for identical seeds); the ring slots it wrote are overwritten by the
        #    first real transition.
        size, head, n_env = self.memory.size, self.memory.head, self.env.n_step
        saved = (self._sched.clone(), self.memory.head_dev.clone(), self.gen.get_state(), self.env.gen.get_state(),
                 [t.clone() for t in self._env_state()])
        self._transition()
        torch.cuda.synchronize()
        self._sched.copy_(saved[0]); self.memory.head_dev.copy_(saved[1])
        self.gen.set_state(saved[2]); self.env.gen.set_state(saved[3])
        for dst, src in zip(self._env_state(), saved[4]):
            dst.copy_(src)
        torch.cuda.synchronize()
        # 2. capture (records the launches, runs nothing); both generators are registered so that every replay advances them
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)
        g.register_generator_state(self.env.gen)
        n0 = int(self._lib.v2v_launch_count())
        try:
            with torch.cuda.graph(g):
                reward = self._transition()
        except Exception as exc:               # a stack that cannot capture this sequence: keep working, launch by launch
            import warnings
            warnings.warn(f"BatchedAgent: CUDA-graph capture of the transition failed ({exc!r}); falling back to "
                          "launch-by-launch transitions")
            self.use_graph = False
            return None
        finally:
            self.memory.size, self.memory.head, self.env.n_step = size, head, n_env     # host counters the traced Python advanced
        per_replay = int(self._lib.v2v_launch_count()) - n0
        self._graph, self._graph_key = (g, reward, per_replay), key
        return self._graph

    def _env_state(self):
        e = self.env
        return [e.pos, e.dir, e.vel, e.v2v_shadow, e.v2i_shadow, e.V2V_channels_with_fastfading, e.V2I_channels_with_fastfading,
                e.V2I_channels_abs]

    def generate_transitions(self, num_transitions):
        """num_transitions steps of every environment into the replay ring (:409-553); returns rewards [T, E] (device).
        With ``use_graph`` every step is one replay of the captured transition."""
        T = int(num_transitions)
        rewards = torch.empty((T, self.E), dtype=torch.float32, device=self.env.dev)
        graph = self._capture() if self.use_graph and T > 0 else None
        for t in range(T):
            if graph is not None:
                graph[0].replay()
                self.replayed_kernel_launches += graph[2]
                rewards[t].copy_(graph[1])
                self.memory._advance(self.E)
                self.env.n_step += 1
            else:
                rewards[t].copy_(self._transition())
            self._num_step += 1
        return rewards

    def replay(self, indices=None):
        """One replay step (:555-748), identical to ``Agent.replay`` but with device-side index sampling.
        ``indices`` (device int64 tensor of ring slots) replaces the random draw -- reproducible replays."""
        B, N, CH = self.batch_size, self.num_D2D, self.num_CH
        m = self.memory
        if indices is not None:
            idx = indices.to(m.device, dtype=torch.int64)
            B = int(idx.numel())
        elif m.size >= B:
            idx = torch.randperm(m.size, generator=self.gen, device=m.device)[:B]           # without replacement (:258-270)
        else:
            idx = torch.randint(0, m.size, (B,), generator=self.gen, device=m.device)
        g = lambda t: t.index_select(0, idx)
        node, edge, node_, edge_ = g(m.node), g(m.edge), g(m.node_), g(m.edge_)
        im, om, action, reward = g(m.in_mask), g(m.out_mask), g(m.action), g(m.reward)
        brain = self.brain
        p = brain.forward_device(node, edge, in_mask=im)                                    # :664
        p_ = brain.forward_device(node_, edge_, in_mask=im, target=True)                    # :665
        y = td_targets_device(self._lib, p, p_, action, reward, self.gamma)                 # :668-692
        losses = brain.train_step_device(node, edge, im, om, None, y)                       # :728
        return losses, y.mean(dim=(0, 2)), p.mean(dim=(0, 2))

    def train(self, num_episodes, num_train_steps, num_transition=50):
        """Agent.train (:750-910) for E environments in lock step.  Returns (loss [episodes, steps, N], mean reward per
        environment step [episodes, steps]) as host arrays."""
        self.total_steps = num_episodes * num_train_steps * num_transition
        self.num_step = 0
        N = self.num_D2D
        loss = np.zeros((num_episodes, num_train_steps, N))
        rew = np.zeros((num_episodes, num_train_steps))
        for ep in range(num_episodes):
            self.env.new_random_game()                                                      # :810
            for it in range(num_train_steps):
                r = self.generate_transitions(num_transition)                               # :827
                l, _, _ = self.replay()                                                     # :832
                if self.num_step % UPDATE_TARGET_FREQUENCY < num_transition:                # :846-847 (steps advance by num_transition)
                    self.brain.update_target_model()
                loss[ep, it] = l.cpu().numpy()
                rew[ep, it] = float(r.mean())
        return loss, rew
