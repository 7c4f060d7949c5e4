#!/usr/bin/env python
"""BASELINE configs[4]: sweep of the aggregation kernel (and the fused brain where it applies) over graph size N and
batch B on ONE GPU; prints a table with the HBM-roofline fraction per point.  Run under gpurun; copy the output to
profiles/sweep_rNN.txt.   python scripts/sweep.py [--quick]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_numpy, load_peaks, L2_BYTES      # noqa: E402

v2v = importlib.import_module("globecom2020-resourceallocationgnn_b200")
lib = v2v.load_library()
ptr = v2v._lib.ptr
PEAK = load_peaks()["hbm_gbs"]


def time_graph(launch, P, reps=10):
    for i in range(P):
        launch(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(P):
            launch(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (reps * P)


def agg_point(B, N, dtype=torch.float32, sparse=0):
    F = 16
    W = (N + 31) // 32
    es = 4 if dtype == torch.float32 else 2
    bytes_per_graph = 2 * N * F * es + N * W * 4
    set_bytes = B * bytes_per_graph
    P = int(max(2, min(16, -(-2 * L2_BYTES // set_bytes))))
    rng = np.random.default_rng(N)
    nb = min(B, 512)
    _, _, adj = synth_numpy(nb, N, rng, sparse)
    im0, _, _ = v2v.pack_adjacency(torch.from_numpy(adj).cuda())
    im0 = im0.repeat((-(-B // nb), 1, 1))[:B].contiguous()
    sets = [(torch.randn((B, N, F), device="cuda").to(dtype), im0.clone(), torch.empty((B, N, F), device="cuda", dtype=dtype))
            for _ in range(P)]
    dt = 0 if dtype == torch.float32 else 1

    def launch(i, flags):
        H, im, out = sets[i]
        v2v._lib.check(lib.v2v_agg_mask_ex(ptr(H), ptr(im), None, ptr(out), B, N, F, dt, flags, v2v._lib.current_stream()))

    us_dep = time_graph(lambda i: launch(i, 0), P)
    us_ind = time_graph(lambda i: launch(i, 1), P)
    return set_bytes, us_dep, us_ind


def brain_point(B, N, S=2, dtype="f32", per_slot=False):
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=per_slot, max_batch=B, data_parallel=False, seed=1, dtype=dtype)
    rng = np.random.default_rng(B + N)
    node, edge, adj = synth_numpy(B, N, rng)
    nd, ed, ad = (torch.from_numpy(t).cuda() for t in (node, edge, adj))
    im, om, _ = v2v.pack_adjacency(ad)
    q = brain.forward_device(nd, ed, in_mask=im)
    y = q + 1.0
    for _ in range(3):
        brain.train_step_device(nd, ed, im, om, None, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        brain.train_step_device(nd, ed, im, om, None, y)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    info = brain.fused_info(B)
    del brain
    return us, info


def brain_f32_section(quick):
    """The fp32 shared-weight brain alone (``--brain-f32``: re-measures this section after a change of the fused kernel),
    with the backward contractions on the tensor cores (the default) and on the FP32 pipe side by side."""
    lib = v2v.load_library()
    print("\n# brain fwd + Huber + bwd + Adam (2 stages, shared weights, fp32): graphs/s on one GPU; backward contractions on the")
    print("# tensor cores (mma.sync TF32, 3 passes: the default) | on the FP32 pipe (V2V_FUSED_MMA=0)")
    print(f"{'N':>4} {'B':>6} {'us/step':>9} {'graphs/s':>12} {'fp32-pipe us':>13}  path")
    for N in (8, 16, 20, 32, 64):
        for B in ((64, 1024, 8192) if quick else (64, 256, 1024, 4096, 8192, 32768)):
            if N > 32 and B > 8192:
                continue
            lib.v2v_fused_set_mma(1)
            us, info = brain_point(B, N)
            us0 = float("nan")
            if info["capable"]:
                lib.v2v_fused_set_mma(0)
                us0, _ = brain_point(B, N)
                lib.v2v_fused_set_mma(1)
            path = f"fused, {info['graphs_per_tile']} graphs/tile" if info["capable"] else "layer-by-layer kernels"
            print(f"{N:4d} {B:6d} {us:9.1f} {B / us * 1e6:12.0f} {us0:13.1f}  {path}", flush=True)


def main():
    quick = "--quick" in sys.argv
    if "--brain-f32" in sys.argv:
        return brain_f32_section(quick)
    Ns = [8, 16, 20, 32, 64, 128, 256]
    Bs = [64, 1024, 8192, 32768] if quick else [64, 256, 1024, 4096, 8192, 32768]
    print(f"# aggregation kernel sweep, fp32, F=16, reference-dense adjacency (E=N(N-2)), one B200, peak {PEAK} GB/s (measured)")
    print(f"# dep = launches serialized by the dependency wait, ind = independent launches (PDL overlap); bytes = algorithmic")
    print(f"{'N':>4} {'B':>6} {'MB/launch':>10} {'dep_us':>8} {'dep_frac':>8} {'ind_us':>8} {'ind_frac':>8}  path")
    for N in Ns:
        for B in Bs:
            if B * N * 16 * 4 * 2 > 6e9:
                continue
            nbytes, us_dep, us_ind = agg_point(B, N)
            path = "dense predicated, warp tiles" if N <= 20 else ("bit walk, warp per graph" if N < 96 else "bit walk, CTA per graph")
            print(f"{N:4d} {B:6d} {nbytes / 1e6:10.2f} {us_dep:8.2f} {nbytes / us_dep / 1e3 / PEAK:8.3f} {us_ind:8.2f} "
                  f"{nbytes / us_ind / 1e3 / PEAK:8.3f}  {path}", flush=True)
    print("\n# bf16 storage (fp32 accumulate)")
    for B in (1024, 8192, 32768):
        nbytes, us_dep, us_ind = agg_point(B, 20, torch.bfloat16)
        print(f"  20 {B:6d} {nbytes / 1e6:10.2f} {us_dep:8.2f} {nbytes / us_dep / 1e3 / PEAK:8.3f} {us_ind:8.2f} "
              f"{nbytes / us_ind / 1e3 / PEAK:8.3f}  TMA fast path, bf16")
    print("\n# sparse 40-link variant (in-degree 2), fp32")
    for B in (1024, 8192):
        nbytes, us_dep, us_ind = agg_point(B, 20, torch.float32, sparse=2)
        print(f"  20 {B:6d} {nbytes / 1e6:10.2f} {us_dep:8.2f} {nbytes / us_dep / 1e3 / PEAK:8.3f} {us_ind:8.2f} "
              f"{nbytes / us_ind / 1e3 / PEAK:8.3f}  TMA fast path, E=40")
    brain_f32_section(quick)
    print("\n# brain fwd + Huber + bwd + Adam, BASELINE configs[2] form: 3 stages, shared weights, bf16 operands on tcgen05 (csrc/tc_train.cu)")
    print(f"{'N':>4} {'B':>6} {'us/step':>9} {'graphs/s':>12}  path")
    for N in (8, 20, 32):
        for B in ((1024, 8192) if quick else (256, 1024, 4096, 8192, 32768)):
            us, _ = brain_point(B, N, S=3, dtype="bf16")
            print(f"{N:4d} {B:6d} {us:9.1f} {B / us * 1e6:12.0f}  tcgen05 bf16, {128 // N} graphs/tile", flush=True)
    print("\n# the reference's own model: per-slot weights, 3 stages, fp32 (fused one-launch kernel up to N = 8)")
    print(f"{'N':>4} {'B':>6} {'us/step':>9} {'graphs/s':>12}  path")
    for N in (4, 8):
        for B in ((256, 4096) if quick else (1, 64, 256, 512, 4096, 32768)):
            us, info = brain_point(B, N, S=3, per_slot=True)
            path = f"fused per-slot, {info['graphs_per_tile']} graphs/tile" if info["capable"] else "layer-by-layer kernels"
            print(f"{N:4d} {B:6d} {us:9.1f} {B / us * 1e6:12.0f}  {path}", flush=True)


if __name__ == "__main__":
    main()
