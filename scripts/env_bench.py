#!/usr/bin/env python
"""Batched environment: environment steps per second on one GPU next to the CPU restatement.

A "step" = Agent.act (BS_brain.py:366-376: reward on the current channels, renew_positions, renew_channels_fastfading) +
the state packing of the next state (BS_brain.py:389-407, :441-469) for ONE environment.  Run under gpurun; the output is
kept in profiles/env_bench_rNN.txt."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_peaks, cpu_env_baseline                 # noqa: E402  (the CPU restatement is timed by bench.py's cpu_baseline leg)

v2v = importlib.import_module("globecom2020-resourceallocationgnn_b200")
PEAK = load_peaks()


def gpu_point(E, N, RB=4, reps=30):
    env = v2v.BatchedEnviron(E, n_veh=N, n_rb=RB, seed=1001)
    env.new_random_game()
    actions = torch.randint(0, RB, (E, N), device="cuda", dtype=torch.int32)
    # pre-drawn randomness: the generator is library plumbing, the timed region is the environment arithmetic
    u = torch.rand((E, N), device="cuda")
    zv, zi = 3 * torch.randn((E, N, N), device="cuda"), 8 * torch.randn((E, N), device="cuda")
    fv, fi = torch.randn((E, N, N, RB, 2), device="cuda"), torch.randn((E, N, RB, 2), device="cuda")

    def step():
        env.compute_reward_with_channel_selection(actions)
        env.renew_positions(u)
        env.renew_channels_fastfading(zv, zi, fv, fi)
        env.pack_state()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    # the channel kernel alone (the HBM streamer of the step)
    for _ in range(3):
        env.renew_channels_fastfading(zv, zi, fv, fi)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        env.renew_channels_fastfading(zv, zi, fv, fi)
    e1.record()
    torch.cuda.synchronize()
    us_ch = 1e3 * e0.elapsed_time(e1) / reps
    # algorithmic bytes of the channel kernel per environment: per (i, j) shadow in/out + draw + 2 RB fading draws + RB out,
    # per i the V2I equivalents, positions and speeds once
    b = N * N * (4 + 4 + 4 + 8 * RB + 4 * RB) + N * (4 + 4 + 4 + 8 * RB + 4 * RB + 4) + N * 12
    # with the draws generated in the timed region (torch.randn / rand: library RNG)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        env.act(actions); env.pack_state()
    e1.record()
    torch.cuda.synchronize()
    us_rng = 1e3 * e0.elapsed_time(e1) / reps
    return us, us_ch, b * E, us_rng


if __name__ == "__main__":
    print(f"# batched environment, one B200; HBM peak {PEAK['hbm_gbs']} GB/s ({PEAK['source']})")
    print(f"# step = reward + renew_positions + renew_channels_fastfading + pack_state; reference Environment.py measured in the")
    print(f"# build container (SURVEY 8 f4): 0.8 ms per step at N = 4, 9.7 ms at N = 20 (one environment, Python loops)")
    print(f"{'E':>6s} {'N':>3s} {'us/step(all E)':>15s} {'env-steps/s':>14s} {'with RNG':>14s} {'channels us':>12s} {'GB/s':>8s} {'frac':>6s}")
    for E, N in ((64, 4), (1024, 4), (8192, 4), (1024, 20), (8192, 20), (32768, 20)):
        us, us_ch, by, us_rng = gpu_point(E, N)
        gbs = by / (us_ch * 1e-6) / 1e9
        print(f"{E:6d} {N:3d} {us:15.1f} {E / (us * 1e-6):14.3e} {E / (us_rng * 1e-6):14.3e} {us_ch:12.1f} {gbs:8.1f} {gbs / PEAK['hbm_gbs']:6.3f}", flush=True)
    # the whole DQN loop on the device: E environments act epsilon-greedily on the brain's Q, transitions go to the replay
    # ring, one replay step (2 forwards + TD target + fwd/bwd/Adam on `batch` sampled transitions) per 4 environment steps
    class Cfg:
        Batch_Size, Gamma, v2v_weight, v2i_weight = 1024, 0.5, 1.0, 0.1
    for E, N, kw in ((1024, 4, dict(stages=3, per_slot=True)), (1024, 20, dict(stages=2, per_slot=False)), (8192, 20, dict(stages=2, per_slot=False))):
        env = v2v.BatchedEnviron(E, n_veh=N, n_rb=4, seed=3)
        agent = v2v.BatchedAgent(env, Cfg, memory_capacity=1 << 16, seed=4, **kw)
        env.new_random_game()
        agent.total_steps = 10 ** 6
        for _ in range(3):
            agent.generate_transitions(4); agent.replay()
        torch.cuda.synchronize(); t0 = time.perf_counter(); reps = 20
        for _ in range(reps):
            agent.generate_transitions(4); agent.replay()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
        print(f"# DQN loop on the device, E={E} N={N} {kw}: {dt * 1e3:.2f} ms per (4 env steps of all E + 1 replay step of {Cfg.Batch_Size}) = "
              f"{4 * E / dt:.3e} transitions/s with learning", flush=True)
    for E, N in ((256, 4), (64, 20)):
        s = cpu_env_baseline(E, N)
        print(f"# CPU restatement (oracle/env_oracle.py, numpy fp64, vectorised over E={E}, mobility in Python loops), N={N}: "
              f"{s * 1e3:.1f} ms per step of all E = {E / s:.3e} env-steps/s on {os.cpu_count()} host cores")
