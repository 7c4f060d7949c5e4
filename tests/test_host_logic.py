"""Host-side logic that needs no GPU: adjacency recovery, sharding, and the data-parallel gradient
rule (world_size 2 over gloo) checked against the full-batch oracle gradient."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import v2v_oracle as O


def test_adjacency_from_input_shapes(v2v):
    rng = np.random.default_rng(0)
    adj = (rng.random((3, 4, 4)) < 0.5).astype(np.float32)
    A = np.stack([np.kron(a, np.eye(16, dtype=np.float32)) for a in adj])
    got = v2v.adjacency_from_input(torch.from_numpy(A), 4, 16)
    assert torch.equal(got, torch.from_numpy(adj))
    assert v2v.adjacency_from_input(torch.from_numpy(adj), 4, 16) is not None
    with pytest.raises(ValueError):
        v2v.adjacency_from_input(torch.zeros(3, 5, 5), 4, 16)
    with pytest.raises(ValueError):
        v2v.adjacency_from_input(torch.zeros(4, 4), 4, 16)


def test_shard_range_partitions_batch(v2v):
    from importlib import import_module
    par = import_module("globecom2020-resourceallocationgnn_b200.parallel")
    for B, W in [(8192, 8), (1000, 8), (7, 2), (3, 4)]:
        spans = [par.shard_range(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _dp_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    par = import_module("globecom2020-resourceallocationgnn_b200.parallel")
    rng = np.random.default_rng(11)
    d = O.BrainDims(num_d2d=6, stages=2, per_slot=False)
    L = O.init_params(d, rng, bias_scale=0.1)
    B = 10
    node, edge, adj, _ = O.synth_batch(B, 6, rng)
    y = O.brain_forward(d, L, node, edge, adj) + rng.normal(0, 1.5, (B, 6, 4))
    lo, hi = par.shard_range(B, rank, world)
    _, ph, g = O.brain_backward(d, L, node[lo:hi], edge[lo:hi], adj[lo:hi], y[lo:hi])
    flat = torch.from_numpy(np.concatenate([O.flatten_params(g), ph]))
    # the engine's rule: SUM all-reduce of [grads | head losses], then scale by the shard weights
    par.allreduce_gradients_(flat, weight=(hi - lo) / (B / world))
    flat /= world
    _, ph_full, g_full = O.brain_backward(d, L, node, edge, adj, y)
    ref = np.concatenate([O.flatten_params(g_full), ph_full])
    err = np.abs(flat.numpy() - ref).max()
    np.save(os.path.join(tmp, f"err{rank}.npy"), np.array(err))
    dist.destroy_process_group()


def test_data_parallel_gradient_rule_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert float(np.load(tmp_path / f"err{r}.npy")) < 1e-12


def _cfg(v2v, N, S=2, per_slot=0, hidden=(80, 40, 20), F=16):
    c = v2v._lib.BrainConfig()
    c.num_d2d, c.node_dim, c.edge_dim, c.feedback, c.num_ch, c.stages, c.per_slot = N, 9, 4, F, 4, S, per_slot
    c.hidden[0], c.hidden[1], c.hidden[2] = hidden
    c.max_batch, c.dtype = 64, 0
    c.lr, c.beta1, c.beta2, c.eps = 1e-3, 0.5, 0.999, 1e-7
    return c


def test_tensor_core_plan_host_query(v2v):
    """The tcgen05 forward's plan (csrc/tc_forward.cu) is host logic: tile size, shared memory, tensor-memory columns and the
    MMA count per tile follow from the brain's dimensions."""
    import ctypes as C
    lib = v2v.load_library()
    info = (C.c_int32 * 8)()
    for N, S in ((20, 2), (4, 3), (32, 1), (7, 2)):
        assert lib.v2v_tc_plan(C.byref(_cfg(v2v, N, S)), info) == 0
        capable, tg, layers, smem, planes, wimg, cols, mmas = list(info)
        assert capable == 1 and tg == 128 // N and layers == S + 4
        assert smem <= 227 * 1024 and planes == 12 and cols <= 256            # two slots share the 512 TMEM columns
        # per tile: stage 0 contracts 16 (2 k-steps), later stages and the first MLP layer 48 (6), then 80, 48, 32: two MMAs each
        assert mmas == 2 * (2 + 6 * (S - 1) + 6 + 10 + 6 + 4)
        assert wimg == 2 * (16 * 16 + (S - 1) * 48 * 16 + 48 * 80 + 80 * 48 + 48 * 32 + 32 * 16) + 64 * -(-(16 * S + 80 + 48 + 32 + 16) // 64)
    # outside the path: per-slot weights (the reference default), N > 32, layers wider than the tensor-memory regions
    for cfg in (_cfg(v2v, 4, 3, per_slot=1), _cfg(v2v, 40), _cfg(v2v, 20, hidden=(200, 40, 20)), _cfg(v2v, 20, F=12)):
        assert lib.v2v_tc_plan(C.byref(cfg), info) == 0 and info[0] == 0


def test_bf16_training_plan_host_query(v2v):
    """csrc/tc_train.cu's plan is host logic too: steps and MMAs per training tile, shared memory, operand planes."""
    import ctypes as C
    lib = v2v.load_library()
    info = (C.c_int32 * 8)()
    for N, S in ((20, 3), (20, 2), (4, 3), (32, 1)):
        assert lib.v2v_tt_plan(C.byref(_cfg(v2v, N, S)), info) == 0
        capable, tg, steps, mmas, smem, planes, wimg, blocks = list(info)
        assert capable == 1 and tg == 128 // N
        assert steps == 4 * S + 8                      # per stage: combine + aggregate (forward), [dh|dagg] + transposed aggregate (backward); 4 MLP, 3 MLP data gradients, 1 weight gradient
        # forward k-steps of 16: 1, 3 per later stage, 3, 5, 3, 2; data gradient: 1, 2, 3, 5, 1 per later stage; every
        # aggregation is 8 k-steps over the tile's 128 rows (S forward, S backward); weight gradient 4 chains x 8 (three of them issued under the backward epilogues)
        assert mmas == (1 + 3 * (S - 1) + 3 + 5 + 3 + 2) + (1 + 2 + 3 + 5 + (S - 1)) + 16 * S + 32
        assert smem <= 226 * 1024 and blocks == S + 7
        assert planes == (2 + 4 * S + 10 + 5 + 3) + (2 * S + 10 + 5 + 3 + 1 + (1 if (2 * S + 19) % 2 else 2))
        assert wimg == 16 * 16 + (S - 1) * 48 * 16 + 48 * 80 + 80 * 48 + 48 * 32 + 32 * 16
    for cfg in (_cfg(v2v, 4, 3, per_slot=1), _cfg(v2v, 40), _cfg(v2v, 20, S=4), _cfg(v2v, 20, F=8)):
        assert lib.v2v_tt_plan(C.byref(cfg), info) == 0 and info[0] == 0


def test_fused_plan_per_slot_host_query(v2v):
    """The fused FP32 kernel's program for the reference's own model (per-slot weights, N = 4, 3 stages): weights stay in
    global memory, rows are slot-major, so the shared-memory footprint is the arena alone."""
    import ctypes as C
    lib = v2v.load_library()
    info = (C.c_int * 8)()
    assert lib.v2v_fused_plan(C.byref(_cfg(v2v, 4, 3, per_slot=1)), 256, 1, info) == 0
    capable, tg, rows, smem, phases, blocks, bias_slots, table = list(info)
    assert capable == 1 and 1 <= tg <= 64 and smem <= 226 * 1024 and blocks == 0 and bias_slots == 0
    assert lib.v2v_fused_plan(C.byref(_cfg(v2v, 4, 3, per_slot=0)), 256, 1, info) == 0 and info[5] > 0
    assert lib.v2v_fused_plan(C.byref(_cfg(v2v, 20, 3, per_slot=1)), 256, 1, info) == 0 and info[0] == 0     # per-slot beyond N = 8: layered


def test_bench_config_is_identical_in_both_arms(monkeypatch):
    """bench.py: the engine arm and the reference arm describe the workload with the SAME `config` object (the driver
    compares them), and --config c3 selects BASELINE configs[2] (3 stages, bf16; strong scaling = 8192 graphs in total)."""
    import sys
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "8", "--config", "c3", "--scaling", "strong"])
    a = bench.parse()
    assert (a.stages, a.dtype, a.batch) == (3, "bf16", 8192) and bench.per_gpu_batch(a, 8) == 1024
    cfg = bench.config_dict(a, 8)
    assert cfg["global_batch"] == 8192 and cfg["graphs_per_gpu"] == 1024 and cfg["parallelism"] == "dp8"
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.config, a.stages, a.dtype, a.batch, a.scaling) == ("c2", 2, "f32", 1024, "weak")
    c1 = bench.config_dict(a, 1)
    assert c1 == bench.config_dict(a, 1) and "configs[1]" in c1["workload"] and c1["edges_per_graph"] == 360


def test_no_undefined_names_in_entry_points_and_package():
    """A NameError inside a rarely taken branch of bench.py only shows on the GPU box (it did once: a config's JSON
    assembly referred to a variable of another config).  Static check: every name a function reads as a global must be
    defined at module level or be a builtin."""
    import builtins
    import glob
    import symtable
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py", "v2v_gnn_b200.py")]
    files += glob.glob(os.path.join(ROOT, "globecom2020-resourceallocationgnn_b200", "*.py")) + glob.glob(os.path.join(ROOT, "scripts", "*.py"))
    bad = []
    for path in files:
        src = open(path).read()
        top = symtable.symtable(src, path, "exec")
        known = set(top.get_identifiers()) | set(dir(builtins)) | {"__file__", "__name__", "__doc__"}

        def walk(t):
            for c in t.get_children():
                if c.get_type() == "function":
                    for s in c.get_symbols():
                        if s.is_global() and s.is_referenced() and s.get_name() not in known:
                            bad.append((os.path.basename(path), c.get_name(), s.get_name()))
                walk(c)
        walk(top)
    assert not bad, bad
