#!/usr/bin/env python
"""Benchmark of the V2V graph-convolution hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a engine
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the reference's CPU form

One "step" = one fwd + Huber + bwd + Keras-Adam pass of the brain (BS.train_dnn, BS_brain.py:218-223)
over one batch of synthetic V2V graphs.  Default workload at every N (what the driver runs): BASELINE.json
configs[1] per GPU -- batch 1024 x 20-vehicle graphs, 2 GNN stages, fp32 -- i.e. weak scaling (1024 graphs
per GPU).  Prints ONE JSON line (rank 0).

Other BASELINE configs, same JSON contract (their lines are committed under profiles/):
    --config c3 [--scaling strong]   configs[2]: batch 8192 total (strong) or 1024 per GPU (weak), 3-stage GNN,
                                     bf16 operands / fp32 accumulate on the tensor cores (csrc/tc_train.cu)
    --config c1                      configs[0]: the reference's own model (N = 4, per-slot weights, 3 stages), forward
                                     at B = 1 and B = 256 from the reference's dict format, with the CPU lines
    --config c4                      configs[3]: the DQN loop (replay batch 256, N = 4) on the device-resident simulator
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "globecom2020-resourceallocationgnn_b200"
L2_BYTES = 126 * 1024 * 1024
SEED = 1001                      # the reference's training seed (RL_Train_main.py:44)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="graphs per GPU")
    ap.add_argument("--nodes", type=int, default=20)
    ap.add_argument("--stages", type=int, default=2)
    ap.add_argument("--per-slot", type=int, default=0)
    ap.add_argument("--sparse", type=int, default=0, help="in-degree of the sparse variant (0 = reference-dense E=N(N-2))")
    ap.add_argument("--agg-batch", type=int, default=8192, help="graphs of the aggregation roofline point")
    ap.add_argument("--pool", type=int, default=0, help="distinct input batches (0 = enough to exceed L2)")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c1", "c4"],
                    help="c2 = BASELINE configs[1] (default), c3 = configs[2] (bf16, 3 stages), c1 = configs[0], c4 = configs[3]")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --batch is the TOTAL over all GPUs")
    ap.add_argument("--dtype", default=None, choices=["f32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the aggregation roofline point (profiling runs)")
    a = ap.parse_args()
    if a.config == "c3":                              # BASELINE configs[2]
        a.stages = 3
        a.dtype = a.dtype or "bf16"
        if a.scaling == "strong" and a.batch == 1024:
            a.batch = 8192
    a.dtype = a.dtype or "f32"
    return a


def per_gpu_batch(a, world):
    return a.batch // world if a.scaling == "strong" else a.batch


def workload_name(a):
    E = a.nodes * (a.sparse if a.sparse else a.nodes - 2)
    which = {"c2": "configs[1]", "c3": "configs[2]"}.get(a.config, a.config)
    size = f"batch {a.batch} in total" if a.scaling == "strong" else f"batch {a.batch} per GPU"
    prec = ("fp32" if a.dtype == "f32" else
            "bf16 contraction operands / fp32 accumulate (tcgen05), fp32 bias-ReLU-aggregation-loss, fp32 master weights + Adam")
    return (f"BASELINE {which}: {size} x {a.nodes}-vehicle graphs, E={E} directed edges/graph "
            f"({'sparse' if a.sparse else 'reference-dense'}), {a.stages}-stage GNN (last stage linear) + 80-40-20-4 MLP, "
            f"{'per-slot' if a.per_slot else 'shared'} weights, fwd+Huber+bwd+Keras-Adam, {prec}")


def config_dict(a, world):
    """The `config` object of the JSON line: the SAME keys and values in both arms (engine and reference)."""
    B = per_gpu_batch(a, world)
    return {"workload": workload_name(a), "config": a.config, "global_batch": world * B, "graphs_per_gpu": B, "nodes": a.nodes,
            "stages": a.stages, "weights": "per-slot" if a.per_slot else "shared", "dtype": a.dtype,
            "edges_per_graph": a.nodes * (a.sparse if a.sparse else a.nodes - 2), "scaling": a.scaling,
            "parallelism": f"dp{world}" if world > 1 else "single"}


# --------------------------------------------------------------------------- synthetic data
def synth_numpy(B, N, rng, sparse=0, CH=4):
    """Feature distributions measured on the reference simulator."""
    v2v = rng.normal(0.66, 0.44, (B, N, CH))
    v2i = rng.normal(0.54, 0.17, (B, N, CH))
    node = np.concatenate([v2v, v2i, np.full((B, N, 1), 10.0)], -1).astype(np.float32)
    edge = rng.normal(0.92, 0.11, (B, N, CH)).astype(np.float32)
    cols = np.tile(np.arange(N), B)
    rows_b = np.repeat(np.arange(B), N)
    if not sparse:
        dest = (np.arange(N)[None, :] + rng.integers(1, N, (B, N))) % N       # BS_brain.py:441-445
        adj = np.ones((B, N, N), np.float32) - np.eye(N, dtype=np.float32)[None]
        adj[rows_b, dest.ravel(), cols] = 0.0
    else:
        adj = np.zeros((B, N, N), np.float32)
        r = rng.integers(0, N - 1, (B, N))
        for k in range(sparse):
            src = (np.arange(N)[None, :] + 1 + (r + k) % (N - 1)) % N
            adj[rows_b, src.ravel(), cols] = 1.0
    return node, edge, adj


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                                      # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        if self._t:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_run(a, steps, warmup, budget_s, note):
    """The reference's own CPU form (oracle/torch_ref.py: per-slot layer calls, (B,NF)x(B,NF,NF) bmm
    against the dense Kronecker adjacency, autograd, Keras-Adam) on all host cores."""
    import torch
    from oracle import v2v_oracle as O
    from oracle import torch_ref as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(SEED)
    d = O.BrainDims(a.nodes, stages=a.stages, per_slot=bool(a.per_slot))
    layers = O.init_params(d, rng)

    def make(bs):
        node, edge, adj = synth_numpy(bs, a.nodes, rng, a.sparse)
        A = np.kron(adj, np.eye(d.F, dtype=np.float32))                      # BS_brain.py:603
        y = rng.normal(0, 1, (bs, a.nodes, d.CH)).astype(np.float32)
        return [torch.from_numpy(t) for t in (node, edge, A, y)]

    model = T.ReferenceFormCPU(d, layers, dtype=torch.float32, form="reference")
    # size the per-step sample so that the whole run fits the budget
    probe = make(32)
    model.fit_step(*probe)
    t0 = time.perf_counter(); model.fit_step(*probe); t_probe = time.perf_counter() - t0
    per_graph = t_probe / 32
    bs = a.batch
    if steps is None:                     # time-boxed: as many full batches as fit the budget
        steps = int(max(5, min(200, budget_s / max(per_graph * bs, 1e-6))))
    while bs > 16 and per_graph * bs * (steps + warmup) > 1.5 * budget_s:
        bs //= 2
    data = make(bs)
    for _ in range(warmup):
        model.fit_step(*data)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); model.fit_step(*data); times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    val = bs * steps / total
    sample = (f"{steps} steps (+{warmup} warm-up) of {bs} graphs each (of the {a.batch}-graph batch), reference-form "
              f"fwd+bwd+Keras-Adam, torch-CPU fp32, {note}")
    return {"value": val, "unit": "graphs/s", "cores": int(torch.get_num_threads()), "kind": "port", "sample": sample,
            "ms_per_step": 1e3 * total / steps, "median_ms_per_step": 1e3 * float(np.median(times)), "graphs_per_step": bs}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.config in ("c1", "c4"):
        print(json.dumps({"impl": "reference", "unavailable": f"--config {a.config} prints its CPU lines inside the engine arm"}))
        return
    res = cpu_reference_run(a, a.steps, a.warmup, budget_s=150.0,
                            note="oracle port (TF1/Keras cannot run in this image: BASELINE.md section 2)")
    line = {
        "impl": "reference", "metric": "V2V graphs/sec (fwd+bwd)", "value": res["value"], "unit": "graphs/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a, a.gpus), "device": "host CPU",
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- the engine arm
def run_engine_arm(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    v2v = importlib.import_module(PKG)
    lib = v2v.load_library()
    if a.config == "c1":
        return run_c1(a, v2v, lib, dev)
    if a.config == "c4":
        return run_c4(a, v2v, lib, dev)
    N, S, CH = a.nodes, a.stages, 4
    B = per_gpu_batch(a, world)
    # ---- rank-0-only extras run BEFORE the process group exists: the other ranks wait in the (CPU-side) rendezvous
    #      instead of spinning in an NCCL barrier on their GPUs, and the CPU leg does not compete with spinning ranks
    roof = predict = cpu = None
    clocks = ClockSampler(local)
    if rank == 0:
        if not a.no_roofline:
            roof = agg_roofline(v2v, lib, dev, a.agg_batch, N, a.sparse, clocks)
        if world == 1 and not a.no_roofline:
            if a.dtype == "f32":
                predict = predict_point(v2v, dev, a.agg_batch, N, S, a.sparse)
            if not a.no_cpu_baseline:
                cpu = cpu_baseline_bounded(a)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ptr = v2v._lib.ptr
    brain = v2v.BS(N, 3, 1, 16, 1, CH, stages=S, per_slot=bool(a.per_slot), max_batch=max(B, 1), data_parallel=(world > 1),
                   seed=SEED, dtype=a.dtype)
    brain.update_target_model()
    if world > 1:                                              # identical replicas
        for w in (0, 1):
            dist.broadcast(brain._views[w], src=0)

    # ---- rotating pool of device-resident batches, larger than L2 in total
    rng = np.random.default_rng(SEED + rank)
    per_batch = B * N * (9 + 4 + 4) * 4 + 2 * B * N * 4
    R = a.pool if a.pool > 0 else max(4, -(-int(1.15 * L2_BYTES) // per_batch))
    pool = []
    st = v2v._lib.current_stream
    for i in range(R):
        node, edge, adj = synth_numpy(B, N, rng, a.sparse)
        nd, ed, ad = (torch.from_numpy(t).to(dev) for t in (node, edge, adj))
        im, om, binary = v2v.pack_adjacency(ad)
        assert binary
        p = brain.forward_device(nd, ed, in_mask=im)
        pn = brain.forward_device(nd, ed, in_mask=im, target=True)
        act = torch.from_numpy(rng.integers(0, CH, (B, N)).astype(np.int32)).to(dev)
        rew = torch.from_numpy(rng.normal(10.0, 3.0, B).astype(np.float32)).to(dev)
        y = torch.empty_like(p)
        v2v._lib.check(lib.v2v_td_target(ptr(p), ptr(pn), ptr(act), ptr(rew), 0.5, ptr(y), B, N, CH, st()))   # :668-692
        pool.append((nd, ed, im, om, y))
        del ad
    head_loss = torch.zeros(N, device=dev)
    # the host-side batches of the end-to-end leg are prepared up front too, so that nothing but the calls themselves
    # separates the two timed regions (generating them in between leaves the GPU idle for ~0.1 s)
    host = []
    if not a.no_e2e:
        for i in range(8):
            node, edge, adj = synth_numpy(B, N, rng, a.sparse)
            yh = rng.normal(0, 1, (B, N, CH)).astype(np.float32)
            host.append(({"Node_Input": node, "Edge_Input": edge, "Adjacency_Matrix": adj}, {"Decide_Output": yh}))
    torch.cuda.synchronize()

    def step(i):
        nd, ed, im, om, y = pool[i % R]
        brain.train_step_device(nd, ed, im, om, None, y, head_loss=head_loss)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(a.warmup, 3)):
        step(i)
    barrier()
    launches0 = lib.v2v_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks:
        barrier()
        e0.record()
        for i in range(a.steps):
            step(a.warmup + i)
        e1.record()
        barrier()
    launches = lib.v2v_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * a.steps / (ms_total * 1e-3)
    loss_now = float(head_loss.sum().item())
    assert np.isfinite(loss_now), "non-finite loss in the timed region"

    # ---- e2e: the reference-facing call with HOST buffers (BS.train_dnn), copies inside the timed region
    e2e = None
    if not a.no_e2e:
        for i in range(max(a.warmup, 3)):
            brain.train_dnn(host[i % 8][0], host[i % 8][1], B)
        k_e2e = max(20, min(a.steps, 200))
        barrier()
        t0 = time.perf_counter()
        for i in range(k_e2e):
            brain.train_dnn(host[i % 8][0], host[i % 8][1], B)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = B * N * (9 + 4 + CH) * 4 + 2 * B * N * 4          # node, edge, targets + both adjacency bit masks
        d2h = N * 4 + 4
        e2e = {"value": world * B * k_e2e / float(dt.item()), "unit": "graphs/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": k_e2e,
               "host_threads": int(lib.v2v_host_stage_threads()), "host_cores": os.cpu_count(),
               "api": "BS.train_dnn(x_dict, y_dict, B) -> v2v_brain_train_views: numpy fp32 arrays (node, edge, dense "
                      "(B,N,N) adjacency, targets) read in place by the host worker pool (gather into pinned staging, "
                      "adjacency bit-packed on the host), pipelined H2D, fwd+bwd+Adam, D2H of the per-head losses, one "
                      "stream synchronisation per call"}
        if world == 1:
            e2e["reference_format"] = e2e_reference_format(brain, B, N, CH, rng, a.sparse)
        del host
    if world > 1:
        dist.barrier()
    if rank == 0:
        peaks = load_peaks()
        cfg = config_dict(a, world)
        line = {
            "metric": "V2V graphs/sec (fwd+bwd)", "value": value, "unit": "graphs/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": cfg,
            "collective": ("none" if world == 1 else
                           ("one fused kernel per step: per-CTA partial reduction + gradient push to all peers over "
                            "NVLink (cudaIpc peer stores, per-element epoch flags) + rank-ordered sum + Keras-Adam"
                            if brain._comm is not None else
                            "one NCCL sum all-reduce of the flat fp32 gradient per step, then the Adam kernel")),
            "l2": f"rotating pool of {R} distinct device-resident input batches ({R * per_batch / 2**20:.0f} MiB "
                  f"> 126 MiB L2); roofline loop rotates over buffer sets > L2 as well",
            "final_loss": loss_now,
            "arithmetic": ("bf16 contraction operands on tcgen05, fp32 accumulation, fp32 master weights / loss / Adam"
                           if a.dtype == "bf16" else
                           ("fp32 throughout; the backward contractions of the fused kernel run on the tensor cores as "
                            "three TF32 passes per product (hi/lo operand splits, fp32 accumulation: fp32-grade, gradients "
                            "within 3e-6 of the FP32-pipe form), the forward on the FP32 pipe"
                            if lib.v2v_fused_get_mma() == 1 and not a.per_slot else "fp32 throughout, FP32 pipe")),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "peaks": peaks, "predict": predict,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_reference_format(brain, B, N, CH, rng, sparse, steps=6):
    """BS.train_dnn fed EXACTLY what the reference's Agent feeds (BS_brain.py:495-504, :724-728): one fp64 array per node
    slot and the dense Kronecker adjacency kron(Adj, I_F) (:603) -- (B, N*F, N*F) fp64 = 839 MB per batch at B = 1024,
    N = 20, which is why only a few steps over two batches are timed.  The engine samples A[:, ::F, ::F] in place."""
    import torch
    F = 16
    batches = []
    for _ in range(2):
        node, edge, adj = synth_numpy(B, N, rng, sparse)
        x = {"Adjacency_Matrix": np.kron(adj.astype(np.float64), np.eye(F))}
        for k in range(N):
            x[f"D{k + 1}_Node_Input"] = node[:, k].astype(np.float64)
            x[f"D{k + 1}_Edge_Input"] = edge[:, k].astype(np.float64)
            x[f"D{k + 1}_Neighbor_Input"] = np.zeros((B, F))
        yl = {f"D{k + 1}_Decide_Output": rng.normal(0, 1, (B, CH)) for k in range(N)}
        batches.append((x, yl))
    for i in range(2):
        brain.train_dnn(*batches[i], B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        brain.train_dnn(*batches[i % 2], B)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": B * steps / dt, "unit": "graphs/s", "steps": steps, "ms_per_call": 1e3 * dt / steps,
            "host_bytes_read_per_step": int(B * N * N * 8 + B * N * (9 + 4 + F + CH) * 8),
            "caller_array_bytes_per_step": int(batches[0][0]["Adjacency_Matrix"].nbytes),
            "api": "BS.train_dnn with the reference's own feed: D{k}_Node_Input / D{k}_Edge_Input / D{k}_Neighbor_Input "
                   "(B, .) fp64 per slot, Adjacency_Matrix (B, N*F, N*F) fp64 = kron(Adj, I_F), D{k}_Decide_Output targets"}


def predict_point(v2v, dev, B, N, S, sparse):
    """BS.predict on device-resident inputs (forward only), both kernels, CUDA events over 50 launches each."""
    import torch
    rng = np.random.default_rng(SEED + 7)
    brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False, seed=SEED)
    nb = min(B, 2048)
    node, edge, adj = synth_numpy(nb, N, rng, sparse)
    rep = -(-B // nb)
    nd, ed, ad = (torch.from_numpy(np.tile(t, (rep, 1, 1))[:B]).to(dev) for t in (node, edge, adj))
    im, _, _ = v2v.pack_adjacency(ad)
    q = torch.empty((B, N, 4), device=dev)
    out = {"batch": B, "unit": "us per forward of the whole batch"}
    if not brain.tensor_core_info()["capable"]:
        return None
    ref = None
    for name, mode in (("fp32_pipe_fused_us", 0), ("tcgen05_3xtf32_us", 2)):
        brain.set_tensor_core(mode)
        for _ in range(5):
            brain.forward_device(nd, ed, in_mask=im, out=q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            brain.forward_device(nd, ed, in_mask=im, out=q)
        e1.record()
        torch.cuda.synchronize()
        out[name] = 1e3 * e0.elapsed_time(e1) / 50
        if ref is None:
            ref = q.clone()
        else:
            out["max_rel_diff_between_kernels"] = float((q - ref).abs().max() / ref.abs().max())
    out["graphs_per_s_tcgen05"] = B / (out["tcgen05_3xtf32_us"] * 1e-6)
    out["note"] = ("same inputs and weights; the tcgen05 kernel issues every contraction as tcgen05.mma kind::tf32 on hi/lo "
                   "operand splits (fp32-grade products), accumulators in TMEM; one MMA warp + 16 epilogue warps, two tiles in flight per SM; automatic selection uses it from 2 tiles per SM")
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one aggregation launch, from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "agg_traffic_r02.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "agg_traffic_r01.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["dram_bytes_read"] + d["dram_bytes_write"]


def agg_roofline(v2v, lib, dev, B, N, sparse, clocks):
    """Average launch duration of the aggregation kernel at (B, N, F=16, fp32), cold L2: a CUDA graph of P
    launches over P distinct (H, mask, out) sets whose total exceeds L2, replayed between two events."""
    import torch
    F = 16
    rng = np.random.default_rng(SEED)
    bytes_per_graph = 2 * N * F * 4 + N * ((N + 31) // 32) * 4          # read H + write agg + in_mask bits
    set_bytes = B * bytes_per_graph
    P = max(8, -(-2 * L2_BYTES // set_bytes))
    sets = []
    _, _, adj = synth_numpy(min(B, 2048), N, rng, sparse)
    adj = np.tile(adj, (-(-B // adj.shape[0]), 1, 1))[:B]
    im0, _, _ = v2v.pack_adjacency(torch.from_numpy(adj).to(dev))
    for i in range(P):
        H = torch.randn((B, N, F), device=dev)
        sets.append((H, im0.clone(), torch.empty_like(H)))
    ptr = v2v._lib.ptr

    def launch(i, flags):
        H, im, out = sets[i]
        v2v._lib.check(lib.v2v_agg_mask_ex(ptr(H), ptr(im), None, ptr(out), B, N, F, 0, flags, v2v._lib.current_stream()))

    def timed(flags):
        for i in range(P):
            launch(i, flags)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(P):
                launch(i, flags)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with clocks:
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / (reps * P)

    reps = 20
    us_dep = timed(0)            # each launch waits for its predecessor (as inside the layered brain's step)
    us = timed(1)                # V2V_AGG_INDEPENDENT: distinct buffers, launches may overlap head/tail (PDL)
    peaks = load_peaks()
    achieved = set_bytes / (us * 1e-6) / 1e9
    achieved_dep = set_bytes / (us_dep * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "agg_mask_f16_kernel (neighbour aggregation, AggLayer.call)",
            "point": f"B={B} graphs x N={N} nodes x F=16, fp32, E={N * (sparse if sparse else N - 2)} edges/graph",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "frac_dependent": achieved_dep / peaks["hbm_gbs"],
            "regimes": "frac = throughput regime (a stream of launches over distinct buffers, PDL without the dependency wait: "
                       "the head of launch i+1 overlaps the tail of launch i; an upper bound no product path uses as is); "
                       "frac_dependent = every launch waits for the complete drain of its predecessor (how the layered brain "
                       "chains the kernel); both from the same loop, both against the same measured peak",
            "peak_source": peaks["source"], "algorithmic_bytes_per_launch": set_bytes, "bytes_per_graph": bytes_per_graph,
            "avg_launch_us": us, "launches_timed": reps * P,
            "serialized": {"avg_launch_us": us_dep, "achieved": achieved_dep, "frac": achieved_dep / peaks["hbm_gbs"],
                           "note": "same loop with the dependency wait kept: launch i+1 starts its loads only after launch i "
                                   "has drained; at 21.6 MB a plain cudaMemcpyAsync D2D reaches 0.63 of peak this way "
                                   "(profiles/agg_variants_r01.txt)"},
            "method": f"CUDA graph of {P} back-to-back launches over {P} distinct buffer sets ({P * set_bytes / 2**20:.0f} MiB "
                      f"> L2, so every read misses L2), {reps} replays between two CUDA events on the launch stream; "
                      f"average = elapsed / launches",
            "traffic": load_traffic()}


# --------------------------------------------------------------------------- BASELINE configs[0] and configs[3]
def _median_time(fn, reps, warm=5, sync=None):
    for _ in range(warm):
        fn()
    if sync:
        sync()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        if sync:
            sync()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(np.min(ts))


def run_c1(a, v2v, lib, dev):
    """configs[0]: "Sim_Config default scenario (4 V2V pairs), single BS_brain forward": the reference's own model (N = 4,
    per-slot weights, 3 stages, BS_brain.py:121-200), forward only, B = 1 (acting, :336) and B = 256 (Sim_Config.py:15),
    fed the reference's dict (fp64 per-slot arrays + Kronecker adjacency); next to it the reference-form CPU path
    (per-slot layer calls, (B,NF)x(B,NF,NF) bmm, predict in chunks of 32) and the factored CPU path (the fairer line),
    median and min of >= 30 repetitions each."""
    import torch
    from oracle import v2v_oracle as O
    from oracle import torch_ref as T
    N, F, CH, S = 4, 16, 4, 3
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(SEED)
    d = O.BrainDims(N, stages=S, per_slot=True)
    L = O.init_params(d, rng, dtype=np.float32)
    brain = v2v.BS(N, 3, 1, F, 1, CH, stages=S, per_slot=True, max_batch=256, data_parallel=False, seed=SEED)
    brain.set_flat_params(O.flatten_params(L), 0)
    cpu_ref = T.ReferenceFormCPU(d, L, dtype=torch.float32, form="reference")
    cpu_fac = T.ReferenceFormCPU(d, L, dtype=torch.float32, form="factored")
    out = {"metric": "V2V graphs/sec (forward, BS.predict)", "unit": "graphs/s", "n_gpus": 1, "higher_is_better": True,
           "dtype": "f32", "data": "synthetic", "scaling": "weak", "vs_baseline": None,
           "config": {"workload": "BASELINE configs[0]: Sim_Config default scenario, N=4 V2V pairs, per-slot weights, 3-stage GNN "
                                  "+ 80-40-20-4 MLP, single BS.predict forward from the reference's dict format (fp64 per-slot "
                                  "arrays, Kronecker adjacency)", "config": "c1", "nodes": N, "stages": S, "weights": "per-slot"},
           "points": {}}
    launches0 = lib.v2v_launch_count()
    for B in (1, 256):
        node, edge, adj = synth_numpy(B, N, rng)
        A = np.kron(adj.astype(np.float64), np.eye(F))
        x = {"Adjacency_Matrix": A}
        for k in range(N):
            x[f"D{k + 1}_Node_Input"] = node[:, k].astype(np.float64)
            x[f"D{k + 1}_Edge_Input"] = edge[:, k].astype(np.float64)
            x[f"D{k + 1}_Neighbor_Input"] = np.zeros((B, F))
        q = np.stack(brain.predict(x), 1)
        tn, te, tA, ta = (torch.from_numpy(t) for t in (node, edge, A.astype(np.float32), adj))
        q_ref = cpu_ref.predict(tn, te, tA).numpy()
        err = float(np.abs(q - q_ref).max() / np.abs(q_ref).max())
        assert err <= 1e-4, err
        med, mn = _median_time(lambda: brain.predict(x), 100, sync=torch.cuda.synchronize)
        med_r, mn_r = _median_time(lambda: cpu_ref.predict(tn, te, tA), 30)
        med_f, mn_f = _median_time(lambda: cpu_fac.predict(tn, te, ta), 30)
        # device-resident forward (no host copies): CUDA events over 200 launches
        nd, ed, ad = (torch.from_numpy(t).to(dev) for t in (node, edge, adj))
        im, _, _ = v2v.pack_adjacency(ad)
        qd = torch.empty((B, N, CH), device=dev)
        for _ in range(10):
            brain.forward_device(nd, ed, in_mask=im, out=qd)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            brain.forward_device(nd, ed, in_mask=im, out=qd)
        e1.record()
        torch.cuda.synchronize()
        out["points"][f"B={B}"] = {
            "e2e_predict_us_median": 1e6 * med, "e2e_predict_us_min": 1e6 * mn, "e2e_graphs_per_s": B / med,
            "device_forward_us": 1e3 * e0.elapsed_time(e1) / 200, "device_graphs_per_s": B / (1e-3 * e0.elapsed_time(e1) / 200),
            "cpu_reference_form_us_median": 1e6 * med_r, "cpu_reference_form_us_min": 1e6 * mn_r,
            "cpu_reference_form_graphs_per_s": B / med_r,
            "cpu_factored_form_us_median": 1e6 * med_f, "cpu_factored_form_us_min": 1e6 * mn_f,
            "cpu_factored_form_graphs_per_s": B / med_f, "max_rel_err_vs_cpu_reference_form": err}
    p256 = out["points"]["B=256"]
    out.update({"value": p256["device_graphs_per_s"], "ms_per_step": 1e-3 * p256["device_forward_us"], "steps": 200, "warmup": 10,
                "e2e": {"value": p256["e2e_graphs_per_s"], "unit": "graphs/s",
                        "h2d_bytes_per_step": 256 * N * (9 + 4) * 4 + 256 * N * 4, "d2h_bytes_per_step": 256 * N * CH * 4},
                "cpu_baseline": {"value": p256["cpu_reference_form_graphs_per_s"], "unit": "graphs/s", "cores": cores, "kind": "port",
                                 "sample": "median of 30 reference-form predicts of 256 graphs (chunks of 32), torch-CPU fp32"},
                "gpu_launches": int(lib.v2v_launch_count() - launches0)})
    print(json.dumps(out), flush=True)


def run_c4(a, v2v, lib, dev):
    """configs[3]: "full RL_Train_main DQN loop, replay batch 256, 1 x B200 (end-to-end drop-in check)".  The reference loop
    (BS_brain.py:750-910: per training step 50 environment transitions with one B=1 greedy/eps forward each, then replay =
    two B=256 forwards + TD targets + one fit) on the reference's own model (N = 4, per-slot, 3 stages).  /root/reference
    does not exist on the GPU box, so the simulator is this repo's device-resident restatement (csrc/env.cu, pinned to
    recordings of the unmodified Environment.py in tests/test_gpu_env.py); E = 1 environment is the reference's loop
    shape, E = 256 shows what batching the environments buys."""
    import torch

    class Cfg:                                   # RL_Train_main.py:29-36, :59 (gamma 0.5, v2i weight 0.1), replay batch 256
        Batch_Size, Gamma, v2v_weight, v2i_weight = 256, 0.5, 1.0, 0.1
    N = 4
    out = {"metric": "DQN training steps/sec (50 transitions + replay batch 256 per step)", "unit": "train steps/s", "n_gpus": 1,
           "higher_is_better": True, "dtype": "f32", "data": "synthetic (device-resident simulator)", "scaling": "weak",
           "vs_baseline": None,
           "config": {"workload": "BASELINE configs[3]: DQN loop, N=4 V2V pairs, per-slot weights, 3 stages, replay batch 256, "
                                  "50 transitions per training step, target sync every 500 environment steps", "config": "c4"},
           "points": {}}
    launches0 = lib.v2v_launch_count()
    replayed = 0
    for E in (1, 256):
        env = v2v.BatchedEnviron(E, n_veh=N, n_rb=4, seed=SEED)
        agent = v2v.BatchedAgent(env, Cfg, memory_capacity=1 << 16, seed=SEED, stages=3, per_slot=True)
        agent.train(num_episodes=1, num_train_steps=6, num_transition=50)          # warm-up: fills the ring (>= 256 slots)
        torch.cuda.synchronize()
        steps = 20
        t0 = time.perf_counter()
        loss, rew = agent.train(num_episodes=1, num_train_steps=steps, num_transition=50)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert np.isfinite(loss).all() and np.isfinite(rew).all()
        replayed += agent.replayed_kernel_launches
        out["points"][f"E={E}"] = {"train_steps_per_s": steps / dt, "transitions_per_s": steps * 50 * E / dt,
                                   "ms_per_train_step": 1e3 * dt / steps, "final_loss": float(loss[-1, -1].sum()),
                                   "mean_reward": float(rew.mean())}
    p1 = out["points"]["E=1"]
    out.update({"value": p1["train_steps_per_s"], "ms_per_step": p1["ms_per_train_step"], "steps": 20, "warmup": 6,
                "e2e": {"value": p1["train_steps_per_s"], "unit": "train steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 2 * N * 4 + 4, "note": "state never leaves the device; per step the per-head "
                        "losses and the mean reward come back"},
                "reference_loop_cost": "the reference spends 9.7 ms per environment step at N=20 / 0.8 ms at N=4 in its Python "
                                       "simulator alone (SURVEY.md 8f-4), i.e. >= 40 ms per training step before any TF call",
                "transition_loop": "one CUDA-graph replay per environment step (simulator kernels, forward, epsilon-greedy "
                                   "selection, ring write: 13 launches recorded once); the replay step is launched kernel by kernel",
                "gpu_launches": int(lib.v2v_launch_count() - launches0 + replayed)})
    print(json.dumps(out), flush=True)


def cpu_baseline_bounded(a):
    """~10-30 s of host work: the reference-form step on a bounded sample of the same workload, and -- the fairer CPU line
    BASELINE.md section 3 promises -- the factored-form step (adjacency as (B,N,N), packed layers) on the same sample."""
    res = cpu_reference_run(a, None, 3, budget_s=12.0, note="oracle port timed on the GPU box's host cores")
    out = {k: v for k, v in res.items() if k in ("value", "unit", "cores", "kind", "sample")}
    out["median_ms_per_step"] = res["median_ms_per_step"]
    try:
        out["factored_form"] = cpu_factored_run(a, res["graphs_per_step"], budget_s=6.0)
    except Exception as e:                                    # the reported baseline is the reference-form line above
        out["factored_form"] = {"error": repr(e)}
    return out


def cpu_factored_run(a, bs, budget_s):
    """Same fit step with the aggregation in factored form (einsum over the (B,N,N) adjacency instead of the bmm against the
    dense Kronecker operand) and packed [B,N,.] layer calls: what a CPU implementation free of the reference's form costs."""
    import torch
    from oracle import v2v_oracle as O
    from oracle import torch_ref as T
    rng = np.random.default_rng(SEED)
    d = O.BrainDims(a.nodes, stages=a.stages, per_slot=bool(a.per_slot))
    model = T.ReferenceFormCPU(d, O.init_params(d, rng), dtype=torch.float32, form="factored")
    node, edge, adj = synth_numpy(bs, a.nodes, rng, a.sparse)
    y = rng.normal(0, 1, (bs, a.nodes, d.CH)).astype(np.float32)
    data = [torch.from_numpy(t) for t in (node, edge, adj, y)]
    for _ in range(3):
        model.fit_step(*data)
    times, t_start = [], time.perf_counter()
    while len(times) < 30 or (time.perf_counter() - t_start < budget_s and len(times) < 300):
        t0 = time.perf_counter(); model.fit_step(*data); times.append(time.perf_counter() - t0)
    return {"value": bs / float(np.median(times)), "unit": "graphs/s", "median_ms_per_step": 1e3 * float(np.median(times)),
            "min_ms_per_step": 1e3 * float(np.min(times)), "steps": len(times), "graphs_per_step": bs,
            "cores": int(torch.get_num_threads()), "kind": "port (factored form)"}


# --------------------------------------------------------------------------- batched-environment CPU leg
def cpu_env_baseline(E, N, RB=4, budget_s=8.0):
    """cpu_baseline leg of scripts/env_bench.py: seconds per step of E environments for the numpy restatement of the
    simulator (oracle/env_oracle.py) on the host cores -- the oracle is only ever the thing compared against."""
    from oracle import env_oracle as EO
    rng = np.random.default_rng(0)
    pos = rng.uniform(0, 700, (E, N, 2)); vel = rng.integers(10, 16, (E, N)).astype(float)
    direction = rng.integers(0, 4, (E, N)); dest = (np.arange(N)[None] + rng.integers(1, N, (E, N))) % N
    sv, si = rng.normal(0, 3, (E, N, N)), rng.normal(0, 8, (E, N))
    _, _, v2v_ff, v2i_ff, _, v2i_abs = EO.renew_channels(pos, vel, sv, si, rng.normal(0, 3, (E, N, N)), rng.normal(0, 8, (E, N)),
                                                         rng.normal(size=(E, N, N, RB, 2)), rng.normal(size=(E, N, RB, 2)))
    actions = rng.integers(0, RB, (E, N))
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < budget_s:
        EO.compute_reward(actions, dest, v2v_ff, v2i_ff, v2i_abs)
        pos, direction = EO.renew_positions(pos, direction, vel, rng.random((E, N)))
        sv, si, v2v_ff, v2i_ff, _, v2i_abs = EO.renew_channels(pos, vel, sv, si, rng.normal(0, 3, (E, N, N)), rng.normal(0, 8, (E, N)),
                                                               rng.normal(size=(E, N, N, RB, 2)), rng.normal(size=(E, N, RB, 2)))
        EO.pack_state(dest, v2v_ff, v2i_ff)
        n += 1
    return (time.perf_counter() - t0) / n



def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_engine_arm(a)


if __name__ == "__main__":
    main()
