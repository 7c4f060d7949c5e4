"""A small stand-in with the interface of the reference ``Environ`` (Environment.py:178-507) for GPU tests:
/root/reference does not exist on the GPU box, so the DQN loop is exercised there against this numpy toy
(random dB channel gains, one receiver per vehicle, Shannon-rate reward under co-channel interference)."""
import numpy as np


class _Vehicle:
    def __init__(self):
        self.destinations = []
        self.neighbors = []


class SyntheticEnviron:
    def __init__(self, n_veh=4, n_rb=4, seed=0):
        self.n_Veh, self.n_RB, self.n_Neighbor = n_veh, n_rb, 1
        self.V2V_power_dB_List = [23, 10, 5]
        self.fixed_v2v_power_index = 1
        self.sig2 = 10 ** (-114 / 10)
        self.rng = np.random.default_rng(seed)
        self.vehicles = []
        self.new_random_game(n_veh)

    def new_random_game(self, n_Veh=0):
        if n_Veh > 0:
            self.n_Veh = n_Veh
        N = self.n_Veh
        self.vehicles = [_Vehicle() for _ in range(N)]
        for i, v in enumerate(self.vehicles):
            v.destinations = [int((i + self.rng.integers(1, N)) % N)]
        self.V2V_abs = self.rng.normal(100.0, 12.0, (N, N)) + 50 * np.eye(N)
        self.V2I_abs = self.rng.normal(110.0, 8.0, N)
        self.V2I_channels_abs = self.V2I_abs
        self.renew_channels_fastfading()

    def renew_positions(self):
        self.V2V_abs = self.V2V_abs + self.rng.normal(0, 0.3, self.V2V_abs.shape)

    def renew_channels_fastfading(self):
        N, C = self.n_Veh, self.n_RB
        self.V2V_channels_with_fastfading = self.V2V_abs[:, :, None] - self.rng.normal(0, 3.0, (N, N, C))
        self.V2I_channels_with_fastfading = self.V2I_abs[:, None] - self.rng.normal(0, 3.0, (N, C))

    def Compute_Interference(self, actions):
        pass

    def compute_reward_with_channel_selection(self, actions):
        N, C = self.n_Veh, self.n_RB
        p = self.V2V_power_dB_List[self.fixed_v2v_power_index]
        sig = np.zeros((N, 1)); interf = np.zeros((N, 1)) + self.sig2
        v2i_interf = np.zeros(C) + self.sig2
        for i in range(N):
            ch, rx = int(actions[i, 0]), self.vehicles[i].destinations[0]
            sig[i, 0] = 10 ** ((p - self.V2V_channels_with_fastfading[i, rx, ch]) / 10)
            v2i_interf[ch] += 10 ** ((p - self.V2I_channels_with_fastfading[i, ch]) / 10)
            for k in range(N):
                if k != i and int(actions[k, 0]) == ch:
                    interf[i, 0] += 10 ** ((p - self.V2V_channels_with_fastfading[k, rx, ch]) / 10)
        v2v_rate = np.log2(1 + sig / interf)
        m = min(C, N)
        v2i_rate = np.log2(1 + 10 ** ((23 - self.V2I_abs[:m]) / 10) / v2i_interf[:m])
        return v2v_rate, v2i_rate, interf
