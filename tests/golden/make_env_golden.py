"""Golden vectors for the batched environment: drives the UNMODIFIED /root/reference/Environment.py,
records its random draws and its own outputs, and stores everything oracle/env_oracle.py and the CUDA kernels are
checked against.  Run here (the reference is not on the GPU box):   python tests/golden/make_env_golden.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import Environment                                   # noqa: E402
from oracle import state_packing as SP               # noqa: E402

DIRS = {"u": 0, "d": 1, "l": 2, "r": 3}


class Recorder:
    """Wraps RandomGenerate.gauss_* (Environment.py:14-42) and random.uniform so that every draw is kept."""

    def __init__(self):
        self.gauss = []
        self.uniform = []
        RG = Environment.RandomGenerate
        self._orig = (RG.gauss_one_d, RG.gauss_two_d, RG.gauss_three_d, random.uniform)
        rec = self

        def wrap(f):
            def g(self_, *a):
                out = f(self_, *a)
                rec.gauss.append(np.array(out))
                return out
            return g
        RG.gauss_one_d, RG.gauss_two_d, RG.gauss_three_d = wrap(self._orig[0]), wrap(self._orig[1]), wrap(self._orig[2])

        def uni(a, b):
            v = rec._orig[3](a, b)
            rec.uniform.append(v)
            return v
        random.uniform = uni

    def take(self):
        g, u = self.gauss, self.uniform
        self.gauss, self.uniform = [], []
        return g, u


def snapshot(env):
    return dict(pos=np.array([v.position for v in env.vehicles], float), dir=np.array([DIRS[v.direction] for v in env.vehicles]),
                vel=np.array([v.velocity for v in env.vehicles], float), dest=np.array([v.destinations[0] for v in env.vehicles]),
                v2v_shadow=env.V2Vchannels.Shadow.copy(), v2i_shadow=env.V2Ichannels.Shadow.copy())


def channel_case(n_veh, seed, steps=6):
    rec = Recorder()
    env = SP.make_env(n_veh, seed)
    rec.take()
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("pos", "dir", "vel", "dest", "v2v_shadow0", "v2i_shadow0", "z_v2i", "z_v2v", "ff_v2i", "ff_v2v", "v2v_shadow1",
                           "v2i_shadow1", "v2v_ff", "v2i_ff", "v2v_abs", "v2i_abs", "actions", "v2v_rate", "v2i_rate", "interference",
                           "node", "edge", "adj", "pos1", "dir1", "u")}
    for _ in range(steps):
        s0 = snapshot(env)
        # reward on the current channels
        actions = rng.integers(0, env.n_RB, (n_veh, 1))
        v2v_rate, v2i_rate, interf = env.compute_reward_with_channel_selection(actions.copy())
        st, adj, _ = SP.build_state(env, n_veh, env.n_RB)
        for k, v in (("v2v_ff", env.V2V_channels_with_fastfading), ("v2i_ff", env.V2I_channels_with_fastfading), ("v2v_abs", env.V2V_channels_abs),
                     ("v2i_abs", env.V2I_channels_abs), ("actions", actions[:, 0]), ("v2v_rate", v2v_rate[:, 0]), ("v2i_rate", v2i_rate),
                     ("interference", interf), ("node", st[:, :9]), ("edge", st[:, 9:]), ("adj", adj), ("dest", s0["dest"])):
            out[k].append(np.array(v, float) if k not in ("actions", "dest") else np.array(v))
        # then the step: positions, channels (Agent.act, BS_brain.py:366-376)
        rec.take()
        env.renew_positions()
        _, us = rec.take()
        s1 = snapshot(env)
        env.renew_channels_fastfading()
        g, _ = rec.take()
        z_v2i, z_v2v, re_i, im_i, re_v, im_v = g                           # draw order of renew_channel / update_fast_fading
        for k, v in (("pos", s0["pos"]), ("dir", s0["dir"]), ("vel", s0["vel"]), ("pos1", s1["pos"]), ("dir1", s1["dir"]),
                     ("v2v_shadow0", s1["v2v_shadow"]), ("v2i_shadow0", s1["v2i_shadow"]), ("z_v2i", z_v2i), ("z_v2v", z_v2v),
                     ("ff_v2i", np.stack([re_i, im_i], -1)), ("ff_v2v", np.stack([re_v, im_v], -1)),
                     ("v2v_shadow1", env.V2Vchannels.Shadow), ("v2i_shadow1", env.V2Ichannels.Shadow)):
            out[k].append(np.array(v))
        out["u"].append(np.array(us + [np.nan] * (n_veh - len(us)), float))  # the lazily drawn uniforms, in vehicle order
    # the channels AFTER each step are the next iteration's v2v_ff: store the last ones too
    out["v2v_ff_last"] = [np.array(env.V2V_channels_with_fastfading)]
    out["v2i_ff_last"] = [np.array(env.V2I_channels_with_fastfading)]
    return {k: np.stack(v) for k, v in out.items()}


def mobility_case(seed):
    """Vehicles placed just before crossings and borders, all four directions, with uniform() forced to 0 / 1 / random."""
    rec = Recorder()
    env = SP.make_env(8, seed)
    rng = np.random.default_rng(seed)
    up, down, left, right = env.up_lanes, env.down_lanes, env.left_lanes, env.right_lanes
    starts = []
    for lane in left + right:
        starts.append(("u", [up[1], lane - 0.05])); starts.append(("d", [down[2], lane + 0.05]))
    for lane in up + down:
        starts.append(("r", [lane - 0.05, right[1]])); starts.append(("l", [lane + 0.05, left[3]]))
    starts += [("u", [up[0], 1298.95]), ("d", [down[0], 0.05]), ("l", [0.04, left[0]]), ("r", [749.96, right[0]]),
               ("u", [up[3], 600.0]), ("r", [300.0, right[2]])]
    res = {k: [] for k in ("pos", "dir", "vel", "u", "pos1", "dir1")}
    orig_uniform = random.uniform
    for mode in ("turn", "straight", "random"):
        for k0 in range(0, len(starts), 8):
            chunk = (starts[k0:k0 + 8] + starts[:8])[:8]
            for v, (d, p) in zip(env.vehicles, chunk):
                v.direction, v.position, v.velocity = d, list(p), int(rng.integers(10, 16))
            s0 = snapshot(env)
            draws = []
            if mode == "random":
                def uni(a, b):
                    x = float(rng.random()); draws.append(x); return x
            else:
                val = 0.0 if mode == "turn" else 1.0
                def uni(a, b):
                    draws.append(val); return val
            random.uniform = uni
            env.renew_positions()
            random.uniform = orig_uniform
            s1 = snapshot(env)
            for k, v in (("pos", s0["pos"]), ("dir", s0["dir"]), ("vel", s0["vel"]), ("pos1", s1["pos"]), ("dir1", s1["dir"]),
                         ("u", np.array(draws + [np.nan] * (8 - len(draws)), float))):
                res[k].append(v)
    return {k: np.stack(v) for k, v in res.items()}


def neighbor_case(seed):
    rec = Recorder()
    res = {"pos": [], "cand": []}
    for n in (4, 8, 20):
        env = SP.make_env(n, seed + n)
        z = np.array([[complex(c.position[0], c.position[1]) for c in env.vehicles]])
        dist = abs(z.T - z)
        cand = np.stack([np.argsort(dist[:, i])[1:n - 2] for i in range(n)])                # Environment.py:370-374
        np.savez_compressed(os.path.join(os.path.dirname(__file__), f"sim_neighbors_n{n}.npz"),
                            pos=np.array([v.position for v in env.vehicles], float), cand=cand,
                            dest=np.array([v.destinations[0] for v in env.vehicles]))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for n, seed in ((4, 1001), (8, 7), (20, 1001)):
        np.savez_compressed(os.path.join(here, f"sim_steps_n{n}.npz"), **channel_case(n, seed))
    np.savez_compressed(os.path.join(here, "sim_mobility.npz"), **mobility_case(3))
    neighbor_case(11)
    print("written")
