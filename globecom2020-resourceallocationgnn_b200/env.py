"""Batched V2V environment: E independent copies of the reference simulator, stepped on the device.

Mirrors the parts of ``Environment.Environ`` the DQN loop uses (``new_random_game``, ``renew_positions``,
``renew_channels_fastfading``, ``compute_reward_with_channel_selection``; Environment.py:236-458, :495-506) and the
Agent's state packing (BS_brain.py:389-407, :441-469) for E scenarios at once.  All state lives in device tensors; the
arithmetic runs in the hand-written kernels of ``csrc/env.cu`` behind the C-ABI (``v2v_env_*``); torch only supplies the
random draws (the kernels take them as inputs, which is what lets the tests compare against vectors recorded from the
unmodified reference) and the buffers.  ``pack_state()`` returns exactly what ``BS.forward_device`` /
``BS.train_step_device`` consume, so transitions never touch the host.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr

UP = (3.5 / 2, 3.5 / 2 + 3.5, 250 + 3.5 / 2, 250 + 3.5 + 3.5 / 2, 500 + 3.5 / 2, 500 + 3.5 + 3.5 / 2)      # RL_Train_main.py:62-75
DOWN = (250 - 3.5 - 3.5 / 2, 250 - 3.5 / 2, 500 - 3.5 - 3.5 / 2, 500 - 3.5 / 2, 750 - 3.5 - 3.5 / 2, 750 - 3.5 / 2)
LEFT = (3.5 / 2, 3.5 / 2 + 3.5, 433 + 3.5 / 2, 433 + 3.5 + 3.5 / 2, 866 + 3.5 / 2, 866 + 3.5 + 3.5 / 2)
RIGHT = (433 - 3.5 - 3.5 / 2, 433 - 3.5 / 2, 866 - 3.5 - 3.5 / 2, 866 - 3.5 / 2, 1299 - 3.5 - 3.5 / 2, 1299 - 3.5 / 2)
WIDTH, HEIGHT = 750, 1299


class BatchedEnviron:
    """E environments x n_Veh vehicles x n_RB resource blocks (n_Neighbor = 1 as in the reference, Environment.py:207)."""

    def __init__(self, num_env, n_veh=4, n_rb=4, device=None, seed=None):
        if n_veh < 4 or n_veh % 4 or n_veh > 32:
            raise ValueError("n_Veh must be a multiple of 4 in [4, 32] (vehicles are added four at a time, Environment.py:217-231)")
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("the batched environment runs on a CUDA device (no CPU fallback)")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.E, self.n_Veh, self.n_RB, self.n_Neighbor = int(num_env), int(n_veh), int(n_rb), 1
        self.gen = torch.Generator(device=self.dev)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        E, N, RB = self.E, self.n_Veh, self.n_RB
        f = dict(dtype=torch.float32, device=self.dev)
        self.pos = torch.zeros((E, N, 2), **f)
        self.dir = torch.zeros((E, N), dtype=torch.int32, device=self.dev)
        self.vel = torch.zeros((E, N), **f)
        self.dest = torch.zeros((E, N), dtype=torch.int32, device=self.dev)
        self.v2v_shadow = torch.zeros((E, N, N), **f)
        self.v2i_shadow = torch.zeros((E, N), **f)
        self.V2V_channels_with_fastfading = torch.zeros((E, N, N, RB), **f)
        self.V2I_channels_with_fastfading = torch.zeros((E, N, RB), **f)
        self.V2I_channels_abs = torch.zeros((E, N), **f)
        self.n_step = 0
        self._shadow_scale = None

    # ---------------------------------------------------------------- random draws (plumbing)
    def _randn(self, *shape):
        return torch.randn(shape, generator=self.gen, device=self.dev, dtype=torch.float32)

    def _rand(self, *shape):
        return torch.rand(shape, generator=self.gen, device=self.dev, dtype=torch.float32)

    def _randint(self, lo, hi, *shape):
        return torch.randint(lo, hi, shape, generator=self.gen, device=self.dev)

    # ---------------------------------------------------------------- the reference's methods
    def new_random_game(self):
        """Environment.py:495-506 with add_new_vehicles_by_number (:217-234): per group of four vehicles one lane index,
        one vehicle per direction at a uniform position on its lane, speed uniform in {10..15} m/s; fresh shadowing."""
        E, N = self.E, self.n_Veh
        G = N // 4
        lane = self._randint(0, 6, E, G)
        t = lambda v: torch.tensor(v, dtype=torch.float32, device=self.dev)
        ys = self._randint(0, HEIGHT + 1, E, G, 2).float()
        xs = self._randint(0, WIDTH + 1, E, G, 2).float()
        pos = torch.empty((E, G, 4, 2), dtype=torch.float32, device=self.dev)
        pos[:, :, 0, 0], pos[:, :, 0, 1] = t(DOWN)[lane], ys[..., 0]           # 'd'
        pos[:, :, 1, 0], pos[:, :, 1, 1] = t(UP)[lane], ys[..., 1]             # 'u'
        pos[:, :, 2, 0], pos[:, :, 2, 1] = xs[..., 0], t(LEFT)[lane]           # 'l'
        pos[:, :, 3, 0], pos[:, :, 3, 1] = xs[..., 1], t(RIGHT)[lane]          # 'r'
        self.pos.copy_(pos.reshape(E, N, 2))
        self.dir.copy_(torch.tensor([1, 0, 2, 3], dtype=torch.int32, device=self.dev).repeat(G)[None].expand(E, N))
        self.vel.copy_(self._randint(10, 16, E, N).float())
        self.v2v_shadow.copy_(3.0 * self._randn(E, N, N))                       # V2Vchannels.__init__ -> update_shadow([]) (:57, :75-77)
        self.v2i_shadow.copy_(8.0 * self._randn(E, N))                          # V2Ichannels.__init__ (:135, :149-150)
        self.n_step = 0
        self.renew_channels_fastfading()
        self.renew_neighbor()

    def renew_positions(self, u=None):
        """Environment.py:236-345; ``u`` [E, N]: the uniform draw a vehicle uses when it reaches a crossing."""
        u = self._rand(self.E, self.n_Veh) if u is None else u
        _lib.check(self._lib.v2v_env_renew_positions(ptr(self.pos), ptr(self.dir), ptr(self.vel), ptr(u), self.E, self.n_Veh,
                                                     _lib.current_stream()))

    def renew_channels_fastfading(self, z_v2v=None, z_v2i=None, ff_v2v=None, ff_v2i=None):
        """Environment.py:378-404 (path loss, shadowing AR(1), fast fading) for all environments in one launch."""
        E, N, RB = self.E, self.n_Veh, self.n_RB
        if z_v2v is None and z_v2i is None and ff_v2v is None and ff_v2i is None:
            # all four draws from ONE standard-normal launch, the shadowing scales (3 dB V2V :55, 8 dB V2I :132) by one multiply
            n1, n2, n3, n4 = E * N * N, E * N, E * N * N * RB * 2, E * N * RB * 2
            z = self._randn(n1 + n2 + n3 + n4)
            if self._shadow_scale is None:
                self._shadow_scale = torch.cat([torch.full((n1,), 3.0, device=self.dev), torch.full((n2,), 8.0, device=self.dev)])
            z[:n1 + n2].mul_(self._shadow_scale)
            z_v2v, z_v2i, ff_v2v, ff_v2i = z[:n1], z[n1:n1 + n2], z[n1 + n2:n1 + n2 + n3], z[n1 + n2 + n3:]
        else:
            z_v2v = 3.0 * self._randn(E, N, N) if z_v2v is None else z_v2v          # shadow_std of the V2V links (:55)
            z_v2i = 8.0 * self._randn(E, N) if z_v2i is None else z_v2i             # :132
            ff_v2v = self._randn(E, N, N, RB, 2) if ff_v2v is None else ff_v2v
            ff_v2i = self._randn(E, N, RB, 2) if ff_v2i is None else ff_v2i
        _lib.check(self._lib.v2v_env_renew_channels(
            ptr(self.pos), ptr(self.vel), ptr(self.v2v_shadow), ptr(self.v2i_shadow), ptr(z_v2v), ptr(z_v2i), ptr(ff_v2v), ptr(ff_v2i),
            ptr(self.V2V_channels_with_fastfading), ptr(self.V2I_channels_with_fastfading), ptr(self.V2I_channels_abs), E, N, RB,
            _lib.current_stream()))

    def renew_neighbor(self, u=None):
        """Environment.py:360-376: one receiver per vehicle among the other vehicles, the two farthest excluded."""
        u = self._rand(self.E, self.n_Veh) if u is None else u
        _lib.check(self._lib.v2v_env_choose_destinations(ptr(self.pos), ptr(u), ptr(self.dest), self.E, self.n_Veh, _lib.current_stream()))

    def compute_reward_with_channel_selection(self, actions, v2v_weight=1.0, v2i_weight=0.1):
        """Environment.py:406-458 for every environment: (V2V_Rate [E,N], V2I_Rate [E,min(RB,N)], Interference [E,RB],
        reward [E] = v2v_weight * sum V2V + v2i_weight * sum V2I as in BS_brain.py:515-519)."""
        E, N, RB = self.E, self.n_Veh, self.n_RB
        a = actions.reshape(E, N).to(device=self.dev, dtype=torch.int32).contiguous()
        f = dict(dtype=torch.float32, device=self.dev)
        v2v_rate, v2i_rate = torch.empty((E, N), **f), torch.empty((E, min(RB, N)), **f)
        interference, reward = torch.empty((E, RB), **f), torch.empty((E,), **f)
        _lib.check(self._lib.v2v_env_reward(ptr(a), ptr(self.dest), ptr(self.V2V_channels_with_fastfading),
                                            ptr(self.V2I_channels_with_fastfading), ptr(self.V2I_channels_abs), ptr(v2v_rate), ptr(v2i_rate),
                                            ptr(interference), ptr(reward), float(v2v_weight), float(v2i_weight), E, N, RB,
                                            _lib.current_stream()))
        return v2v_rate, v2i_rate, interference, reward

    def act(self, actions, v2v_weight=1.0, v2i_weight=0.1):
        """Agent.act (BS_brain.py:366-376): reward on the current channels, then move and renew the channels."""
        out = self.compute_reward_with_channel_selection(actions, v2v_weight, v2i_weight)
        self.renew_positions()
        self.renew_channels_fastfading()
        self.n_step += 1
        return out

    def pack_state(self, dense_adj=False):
        """Agent.get_state + packing (BS_brain.py:389-407, :441-469): node [E,N,2RB+1], edge [E,N,RB], the adjacency as the
        two bit-mask orientations the brain consumes (and optionally the dense [E,N,N] matrix)."""
        E, N, RB = self.E, self.n_Veh, self.n_RB
        f = dict(dtype=torch.float32, device=self.dev)
        node, edge = torch.empty((E, N, 2 * RB + 1), **f), torch.empty((E, N, RB), **f)
        im = torch.empty((E, N, 1), dtype=torch.int32, device=self.dev)
        om = torch.empty((E, N, 1), dtype=torch.int32, device=self.dev)
        adj = torch.empty((E, N, N), **f) if dense_adj else None
        _lib.check(self._lib.v2v_env_pack_state(ptr(self.dest), ptr(self.V2V_channels_with_fastfading), ptr(self.V2I_channels_with_fastfading),
                                                ptr(node), ptr(edge), ptr(im), ptr(om), ptr(adj), E, N, RB, _lib.current_stream()))
        return (node, edge, im, om, adj) if dense_adj else (node, edge, im, om)
