"""Generate the committed golden vectors under tests/golden/ (run in the build container).

    python tests/golden/make_golden.py

Uses (a) the unmodified reference simulator /root/reference/Environment.py to produce
real N=4 states through the restated Agent packing (oracle/state_packing.py), and
(b) the fp64 NumPy oracle (oracle/v2v_oracle.py) for every output.  Inputs and
weights are rounded to fp32 *before* the oracle runs so that the GPU path is fed
bit-identical numbers.  The reference's own network (Keras/TF1) cannot run here:
the outputs are oracle outputs (parity unpinned by the reference, see the oracle header).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import v2v_oracle as O          # noqa: E402
from oracle import state_packing as SP      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def f32(x):
    return np.asarray(x, np.float32).astype(np.float64)


def run_case(name, dims, node, edge, adj, rng, n_adam=3, gamma=0.5):
    node, edge, adj = f32(node), f32(edge), f32(adj)
    layers = O.init_params(dims, rng, bias_scale=0.05)
    for l in layers:
        l['W'], l['b'] = f32(l['W']), f32(l['b'])
    tgt_layers = O.init_params(dims, rng, bias_scale=0.05)
    for l in tgt_layers:
        l['W'], l['b'] = f32(l['W']), f32(l['b'])
    B, N = node.shape[:2]
    q, tape = O.brain_forward(dims, layers, node, edge, adj, keep=True)
    q_t = O.brain_forward(dims, tgt_layers, node, edge, adj)
    actions = rng.integers(0, dims.CH, (B, N))
    rewards = f32(rng.normal(10.0, 3.0, B))
    y = f32(O.td_targets(q, q_t, actions, rewards, gamma))
    loss, per_head, grads = O.brain_backward(dims, layers, node, edge, adj, y)
    flat = O.flatten_params(layers)
    gflat = O.flatten_params(grads)
    # a few Keras-Adam steps on the same batch
    p, m, v = flat.copy(), np.zeros_like(flat), np.zeros_like(flat)
    losses = []
    for t in range(1, n_adam + 1):
        lay = O.unflatten_params(dims, p)
        l_t, _, g_t = O.brain_backward(dims, lay, node, edge, adj, y)
        losses.append(l_t)
        p, m, v = O.keras_adam_step(p, O.flatten_params(g_t), m, v, t)
    # intermediates of the first aggregation for kernel-level checks
    S = dims.S
    h0 = np.maximum(tape['pre'][0], 0) if S > 1 else tape['pre'][0]
    agg0 = O.agg_factored(h0, adj)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        N=N, S=S, per_slot=int(dims.per_slot), F=dims.F, CH=dims.CH,
        node=node.astype(np.float32), edge=edge.astype(np.float32), adj=adj.astype(np.float32),
        params=flat.astype(np.float32), target_params=O.flatten_params(tgt_layers).astype(np.float32),
        q=q, q_target=q_t, actions=actions.astype(np.int32), rewards=rewards.astype(np.float32), gamma=gamma,
        y=y.astype(np.float32), loss=loss, per_head=per_head, grads=gflat,
        h0=h0, agg0=agg0, params_after_adam=p, losses_adam=np.array(losses))
    print(f"{name}: B={B} N={N} S={S} per_slot={dims.per_slot} params={flat.size} loss={loss:.6f}")


def env_states(n_veh, B, seed):
    """B consecutive states of the unmodified reference simulator under random actions."""
    env = SP.make_env(n_veh, seed)
    rng = np.random.default_rng(seed)
    nodes, edges, adjs = [], [], []
    for step in range(B):
        state, adj, _ = SP.build_state(env, n_veh, env.n_RB)
        nodes.append(state[:, :2 * env.n_RB + 1])
        edges.append(state[:, 2 * env.n_RB + 1:])
        adjs.append(adj)
        actions = rng.integers(0, env.n_RB, (n_veh, 1))
        env.compute_reward_with_channel_selection(actions.copy())     # BS_brain.py:366-376 (act)
        env.renew_positions()
        env.renew_channels_fastfading()
        env.Compute_Interference(actions)
        if step % 3 == 2:
            env.renew_neighbor()                                      # new destinations -> new adjacency
    return np.stack(nodes), np.stack(edges), np.stack(adjs)


def main():
    # C1: the reference shape, real simulator states (N=4, per-slot weights, 3 stages)
    node, edge, adj = env_states(4, 8, 1001)
    run_case("ref_n4_env", O.BrainDims(4, stages=3, per_slot=True), node, edge, adj, np.random.default_rng(1001))
    # N=20 from the real simulator too (shared weights, 3 stages)
    node, edge, adj = env_states(20, 6, 1002)
    run_case("env_n20_shared_s3", O.BrainDims(20, stages=3, per_slot=False), node, edge, adj,
             np.random.default_rng(1002))
    # C2 shape: synthetic N=20, 2 stages, shared weights, dense E=360
    rng = np.random.default_rng(1003)
    node, edge, adj, _ = O.synth_batch(16, 20, rng)
    run_case("synth_n20_shared_s2", O.BrainDims(20, stages=2, per_slot=False), node, edge, adj, rng)
    # sparse 40-link variant (in-degree 2), 3 stages
    rng = np.random.default_rng(1004)
    node, edge, adj, _ = O.synth_batch(8, 20, rng, sparse_in_degree=2)
    run_case("synth_n20_sparse40_s3", O.BrainDims(20, stages=3, per_slot=False), node, edge, adj, rng)
    # per-slot weights at N=8 (grouped path, N not a multiple of 4 words etc.)
    rng = np.random.default_rng(1005)
    node, edge, adj, _ = O.synth_batch(5, 7, rng)
    run_case("synth_n7_perslot_s3", O.BrainDims(7, stages=3, per_slot=True), node, edge, adj, rng)


if __name__ == "__main__":
    main()
