"""CPU restatement of the reference simulator's per-step arithmetic, vectorised over E independent environments.

TEST INFRASTRUCTURE ONLY (same import rules as oracle/v2v_oracle.py): only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it.  Every function cites the lines of /root/reference/Environment.py (or
BS_brain.py) it follows; the reference runs in this image, so the restatement is PINNED: tests/golden/make_env_golden.py
drives the unmodified Environment.Environ, copies its state into these functions and stores inputs and the reference's
own outputs in tests/golden/sim_*.npz (tests/test_env_oracle.py checks them to 1e-12, fp64).

Randomness is INJECTED: the reference fills its Gaussian arrays element by element from Python's `random.gauss`
(Environment.py:14-42) and draws `random.uniform` lazily at lane crossings (:251, :259, ...); here every function takes
the draws as arrays, so the arithmetic can be compared exactly and the device path can use any generator.
"""
from __future__ import annotations

import numpy as np

# ---- constants of the reference -------------------------------------------------------------------------------
UP = np.array([3.5 / 2, 3.5 / 2 + 3.5, 250 + 3.5 / 2, 250 + 3.5 + 3.5 / 2, 500 + 3.5 / 2, 500 + 3.5 + 3.5 / 2])       # RL_Train_main.py:62-75
DOWN = np.array([250 - 3.5 - 3.5 / 2, 250 - 3.5 / 2, 500 - 3.5 - 3.5 / 2, 500 - 3.5 / 2, 750 - 3.5 - 3.5 / 2, 750 - 3.5 / 2])
LEFT = np.array([3.5 / 2, 3.5 / 2 + 3.5, 433 + 3.5 / 2, 433 + 3.5 + 3.5 / 2, 866 + 3.5 / 2, 866 + 3.5 + 3.5 / 2])
RIGHT = np.array([433 - 3.5 - 3.5 / 2, 433 - 3.5 / 2, 866 - 3.5 - 3.5 / 2, 866 - 3.5 / 2, 1299 - 3.5 - 3.5 / 2, 1299 - 3.5 / 2])
WIDTH, HEIGHT = 750.0, 1299.0
TIMESTEP = 0.01                       # Environment.py:183
DIR_U, DIR_D, DIR_L, DIR_R = 0, 1, 2, 3
V2V_POWER_DB_LIST = (23.0, 10.0, 5.0)  # :194
FIXED_POWER_INDEX = 1                 # :195
V2I_POWER_DB = 23.0                   # :193
SIG2 = 10.0 ** (-114.0 / 10.0)        # :196, :201
BS_ANT_GAIN, BS_NOISE_FIGURE, VEH_ANT_GAIN, VEH_NOISE_FIGURE = 8.0, 5.0, 3.0, 9.0     # :197-200
BS_POSITION = (750.0 / 2, 1299.0 / 2)  # :131
CONST_A, CONST_B = 80.0, 60.0         # BS_brain.py:392-393


# ---- large-scale fading ---------------------------------------------------------------------------------------
def v2v_path_loss(pos):
    """V2Vchannels.get_path_loss / update_pathloss (Environment.py:63-68, :93-120).  pos [E,N,2] -> [E,N,N]."""
    h_bs = h_ms = 1.5
    fc = 2.0
    d1 = np.abs(pos[:, :, None, 0] - pos[:, None, :, 0])
    d2 = np.abs(pos[:, :, None, 1] - pos[:, None, :, 1])
    d = np.hypot(d1, d2) + 0.001
    d_bp = 4 * (h_bs - 1) * (h_ms - 1) * fc * (10 ** 9) / (3 * 10 ** 8)

    def pl_los(x):
        near = 22.7 * np.log10(3.0) + 41 + 20 * np.log10(fc / 5)
        mid = 22.7 * np.log10(x) + 41 + 20 * np.log10(fc / 5)
        far = 40.0 * np.log10(x) + 9.45 - 17.3 * np.log10(h_bs) - 17.3 * np.log10(h_ms) + 2.7 * np.log10(fc / 5)
        return np.where(x <= 3, near, np.where(x < d_bp, mid, far))

    def pl_nlos(da, db):
        n_j = np.maximum(2.8 - 0.0024 * db, 1.84)
        with np.errstate(divide="ignore", invalid="ignore"):
            return pl_los(da) + 20 - 12.5 * n_j + 10 * n_j * np.log10(db) + 3 * np.log10(fc / 5)

    with np.errstate(divide="ignore", invalid="ignore"):
        los = pl_los(d)
        nlos = np.minimum(pl_nlos(d1, d2), pl_nlos(d2, d1))
    return np.where(np.minimum(d1, d2) < 7, los, nlos)


def v2i_path_loss(pos):
    """V2Ichannels.update_pathloss (Environment.py:140-146).  pos [E,N,2] -> [E,N]."""
    h_bs, h_ms = 25.0, 1.5
    d1 = np.abs(pos[..., 0] - BS_POSITION[0])
    d2 = np.abs(pos[..., 1] - BS_POSITION[1])
    distance = np.hypot(d1, d2)
    return 128.1 + 37.6 * np.log10(np.sqrt(distance ** 2 + (h_bs - h_ms) ** 2) / 1000)


def v2v_shadow_update(shadow, delta, z):
    """V2Vchannels.update_shadow (Environment.py:70-83): shadow [E,N,N], delta [E,N] (metres moved), z ~ N(0, 3) draws."""
    dd = delta[:, :, None] + delta[:, None, :]
    return np.exp(-1 * (dd / 10.0)) * shadow + np.sqrt(1 - np.exp(-2 * (dd / 10.0))) * z


def v2i_shadow_update(shadow, delta, z):
    """V2Ichannels.update_shadow (Environment.py:148-156): shadow [E,N], delta [E,N], z ~ N(0, 8) draws."""
    return np.exp(-1 * (delta / 50.0)) * shadow + np.sqrt(1 - np.exp(-2 * (delta / 50.0))) * z


def fast_fading(real, imag):
    """update_fast_fading (Environment.py:85-91, :158-165): 20 log10 |(re + j im) / sqrt(2)|."""
    h = 1 / np.sqrt(2) * (real + 1j * imag)
    return 20 * np.log10(np.abs(h))


def renew_channels(pos, vel, v2v_shadow, v2i_shadow, z_v2v, z_v2i, ff_v2v, ff_v2i):
    """renew_channel + renew_channels_fastfading (Environment.py:378-404).

    ff_v2v [E,N,N,RB,2] / ff_v2i [E,N,RB,2]: the standard-normal real and imaginary draws.  Returns the new shadows and
    V2V_channels_with_fastfading [E,N,N,RB], V2I_channels_with_fastfading [E,N,RB], V2V_channels_abs, V2I_channels_abs."""
    N = pos.shape[1]
    delta = 0.002 * vel                                                      # :386
    v2i_shadow = v2i_shadow_update(v2i_shadow, delta, z_v2i)
    v2v_shadow = v2v_shadow_update(v2v_shadow, delta, z_v2v)
    v2v_abs = v2v_path_loss(pos) + v2v_shadow + 50 * np.identity(N)[None]   # :389-390
    v2i_abs = v2i_path_loss(pos) + v2i_shadow                               # :391
    v2v_ff = v2v_abs[..., None] - fast_fading(ff_v2v[..., 0], ff_v2v[..., 1])   # :399-401
    v2i_ff = v2i_abs[..., None] - fast_fading(ff_v2i[..., 0], ff_v2i[..., 1])   # :402-404
    return v2v_shadow, v2i_shadow, v2v_ff, v2i_ff, v2v_abs, v2i_abs


# ---- rewards --------------------------------------------------------------------------------------------------
def compute_reward(actions, dest, v2v_ff, v2i_ff, v2i_abs):
    """compute_reward_with_channel_selection (Environment.py:406-458) with n_Neighbor = 1 (:207), all links active (:506)
    and the fixed power index (:195).  actions, dest [E,N] ints.  Returns V2V_Rate [E,N], V2I_Rate [E,min(RB,N)],
    Interference [E,RB] (the V2I interference before the noise floor is added, as the reference returns it)."""
    E, N = actions.shape
    RB = v2i_ff.shape[-1]
    P = V2V_POWER_DB_LIST[FIXED_POWER_INDEX]
    e_idx = np.arange(E)[:, None]
    i_idx = np.arange(N)[None, :]
    # V2I interference per resource block (:413-420)
    contrib = 10 ** ((P - v2i_ff[e_idx, i_idx, actions] + VEH_ANT_GAIN + BS_ANT_GAIN - BS_NOISE_FIGURE) / 10)
    interference = np.zeros((E, RB))
    for rb in range(RB):
        interference[:, rb] = np.where(actions == rb, contrib, 0.0).sum(1)
    v2i_interference = interference + SIG2
    # V2V links: transmitter j -> receiver dest[j] on channel actions[j] (:425-451)
    recv = dest
    gain = lambda tx, rx, ch: 10 ** ((P - v2v_ff[e_idx, tx, rx, ch] + 2 * VEH_ANT_GAIN - VEH_NOISE_FIGURE) / 10)
    signal = gain(i_idx, recv, actions)
    v2v_interf = np.zeros((E, N))
    # the V2I link of resource block c is transmitted by vehicle c (:436-440)
    c = actions
    v2i_tx = 10 ** ((V2I_POWER_DB - v2v_ff[e_idx, np.minimum(c, N - 1), recv, c] + 2 * VEH_ANT_GAIN - VEH_NOISE_FIGURE) / 10)
    v2v_interf += np.where(c < N, v2i_tx, 0.0)
    for k in range(N):                                                       # every other link on the same channel (:442-451)
        same = (actions[:, k:k + 1] == actions) & (i_idx != k)
        g = 10 ** ((P - v2v_ff[e_idx, k, recv, actions] + 2 * VEH_ANT_GAIN - VEH_NOISE_FIGURE) / 10)
        v2v_interf += np.where(same, g, 0.0)
    v2v_rate = np.log2(1 + signal / (v2v_interf + SIG2))                   # :452-453
    m = min(RB, N)
    v2i_signals = V2I_POWER_DB - v2i_abs[:, :m] + VEH_ANT_GAIN + BS_ANT_GAIN - BS_NOISE_FIGURE   # :454-455
    v2i_rate = np.log2(1 + 10 ** (v2i_signals / 10) / v2i_interference[:, :m])   # :456
    return v2v_rate, v2i_rate, interference


# ---- state packing (BS_brain.py:389-407, :441-469) --------------------------------------------------------------
def pack_state(dest, v2v_ff, v2i_ff):
    """Per node [V2V gain x RB | V2I gain x RB | power] and edge x RB, plus Adj[n][m] = 1 - I with Adj[dest[m]][m] = 0."""
    E, N = dest.shape
    e_idx = np.arange(E)[:, None]
    i_idx = np.arange(N)[None, :]
    v2v = (v2v_ff[e_idx, i_idx, dest, :] - CONST_A) / CONST_B
    v2i = (v2i_ff - CONST_A) / CONST_B
    col_sum = v2v_ff.sum(1)                                                  # [E, rx, RB]: sum over transmitters
    edge = (((col_sum[e_idx, dest, :] - v2v_ff[e_idx, dest, dest, :]) - (N - 1) * CONST_A) / CONST_B - v2v) / (N - 2)
    power = np.full((E, N, 1), V2V_POWER_DB_LIST[FIXED_POWER_INDEX])
    node = np.concatenate([v2v, v2i, power], -1)
    adj = np.ones((E, N, N)) - np.eye(N)[None]
    adj[e_idx, dest, i_idx] = 0.0
    return node, edge, adj


# ---- mobility (Environment.py:236-345) ----------------------------------------------------------------------------
def renew_positions(pos, direction, vel, u):
    """One 10 ms step of every vehicle.  u [E,N]: the uniform draw a vehicle uses if it reaches a crossing this step
    (the reference draws lazily; lanes of the two perpendicular directions are >= 3.5 m apart and a step is <= 0.15 m,
    so at most one crossing, hence one draw, per vehicle per step).  Returns new (pos, direction)."""
    pos = pos.copy()
    direction = direction.copy()
    E, N = direction.shape
    dd = vel * TIMESTEP
    for e in range(E):
        for i in range(N):
            x, y = pos[e, i]
            d = direction[e, i]
            step = dd[e, i]
            turned = False
            if d == DIR_U:
                for lanes, sign, nd in ((LEFT, -1.0, DIR_L), (RIGHT, +1.0, DIR_R)):
                    for lane in lanes:
                        if y <= lane and y + step >= lane:
                            if u[e, i] < 0.4:
                                x = x - (step - (lane - y)) if sign < 0 else x + (step + (lane - y))
                                y, d, turned = lane, nd, True
                            break
                    if turned:
                        break
                if not turned:
                    y += step
            elif d == DIR_D:
                for lanes, sign, nd in ((LEFT, -1.0, DIR_L), (RIGHT, +1.0, DIR_R)):
                    for lane in lanes:
                        if y >= lane and y - step <= lane:
                            if u[e, i] < 0.4:
                                x = x - (step - (y - lane)) if sign < 0 else x + (step + (y - lane))
                                y, d, turned = lane, nd, True
                            break
                    if turned:
                        break
                if not turned:
                    y -= step
            elif d == DIR_R:
                for lanes, sign, nd in ((UP, +1.0, DIR_U), (DOWN, -1.0, DIR_D)):
                    for lane in lanes:
                        if x <= lane and x + step >= lane:
                            if u[e, i] < 0.4:
                                y = y + (step - (lane - x)) if sign > 0 else y - (step - (lane - x))
                                x, d, turned = lane, nd, True
                            break
                    if turned:
                        break
                if not turned:
                    x += step
            else:
                for lanes, sign, nd in ((UP, +1.0, DIR_U), (DOWN, -1.0, DIR_D)):
                    for lane in lanes:
                        if x >= lane and x - step <= lane:
                            if u[e, i] < 0.4:
                                y = y + (step - (x - lane)) if sign > 0 else y - (step - (x - lane))
                                x, d, turned = lane, nd, True
                            break
                    if turned:
                        break
                if not turned:
                    x -= step
            if x < 0 or y < 0 or x > WIDTH or y > HEIGHT:                   # leaves the map: re-enter on the border lane (:323-343)
                if d == DIR_U:
                    d, y = DIR_R, RIGHT[-1]
                elif d == DIR_D:
                    d, y = DIR_L, LEFT[0]
                elif d == DIR_L:
                    d, x = DIR_U, UP[0]
                else:
                    d, x = DIR_D, DOWN[-1]
            pos[e, i] = (x, y)
            direction[e, i] = d
    return pos, direction


def crossing(pos, direction, vel):
    """Which vehicles reach a perpendicular lane this step (and therefore consume a uniform draw in the reference)."""
    E, N = direction.shape
    dd = vel * TIMESTEP
    out = np.zeros((E, N), bool)
    for e in range(E):
        for i in range(N):
            x, y = pos[e, i]
            d, step = direction[e, i], dd[e, i]
            if d == DIR_U:
                out[e, i] = any(y <= l and y + step >= l for l in np.concatenate([LEFT, RIGHT]))
            elif d == DIR_D:
                out[e, i] = any(y >= l and y - step <= l for l in np.concatenate([LEFT, RIGHT]))
            elif d == DIR_R:
                out[e, i] = any(x <= l and x + step >= l for l in np.concatenate([UP, DOWN]))
            else:
                out[e, i] = any(x >= l and x - step <= l for l in np.concatenate([UP, DOWN]))
    return out


def destination_candidates(pos):
    """renew_neighbor (Environment.py:360-376): for vehicle i the candidate receivers are sort_idx[1 : N-2], i.e. all other
    vehicles except the two farthest, in order of distance.  Returns [E,N,N-3] indices."""
    z = pos[..., 0] + 1j * pos[..., 1]
    dist = np.abs(z[:, :, None] - z[:, None, :])                            # Distance[:, i] = distances to vehicle i
    order = np.argsort(dist, axis=1, kind="stable")                         # sorted along the first vehicle axis, per column i
    N = pos.shape[1]
    return np.transpose(order[:, 1:N - 2, :], (0, 2, 1))                   # [E, i, rank]


def choose_destinations(pos, u):
    """One receiver per vehicle: candidate floor(u * (N-3)) of destination_candidates (the reference uses random.sample)."""
    cand = destination_candidates(pos)
    k = np.minimum((u * cand.shape[-1]).astype(np.int64), cand.shape[-1] - 1)
    return np.take_along_axis(cand, k[..., None], -1)[..., 0]
