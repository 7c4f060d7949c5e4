"""Host-side mirror of the reference's ``BS`` brain (BS_brain.py:90-239).

Same constructor, attributes and methods the reference's ``Agent`` calls
(``predict`` / ``predict_one_step`` / ``train_dnn`` / ``update_target_model`` and
``model`` / ``target_model`` with ``get_weights`` / ``set_weights`` / ``save_weights`` /
``load_weights``), same input/label dictionary keys (:495-504, :724-725) and history
keys (:836).  All arithmetic runs in the engine behind the C-ABI
(include/v2v_gnn.h); this file only adapts the reference's dict-of-numpy-arrays
calling convention to the engine's packed buffers and, for multi-GPU data
parallelism, places one NCCL all-reduce of the flat gradient between the backward
and the optimiser.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import BrainConfig, ptr
from .layers import _device


class History:
    """What ``Model.fit`` returns; the agent reads ``.history[...][0]`` (BS_brain.py:836-837)."""

    def __init__(self):
        self.history = {}
        self.epoch = []
        self.params = {}


class _DevBuf:
    """torch view over engine-owned device memory (no copy) via __cuda_array_interface__."""

    def __init__(self, address: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (address, False), "version": 3,
                                         "strides": None}


class _ModelHandle:
    """Stands in for the ``keras.models.Model`` objects ``BS.model`` / ``BS.target_model``."""

    def __init__(self, brain, which: int):
        self._brain = brain
        self._which = which       # 0 online, 1 target

    # Model.predict(x) -> list of N arrays (B, CH)   (BS_brain.py:229-231)
    def predict(self, x, batch_size=32, verbose=0):
        return self._brain.predict(x, target=bool(self._which))

    def fit(self, x, y, batch_size=None, epochs=1, verbose=0, shuffle=True):
        if self._which != 0:
            raise RuntimeError("only the online model is compiled for training (BS_brain.py:212-214)")
        return self._brain._fit(x, y, batch_size, epochs, shuffle)

    def get_weights(self):
        return self._brain._get_weight_list(self._which)

    def set_weights(self, weights):
        self._brain._set_weight_list(self._which, weights)

    def count_params(self):
        return int(self._brain.param_count)

    def save_weights(self, filepath, overwrite=True):
        """Keras writes HDF5 (BS_brain.py:863-870); h5py is not part of this stack, so the same
        ordered weight list is stored as ``.npz`` (keys w000, w001, ...)."""
        ws = self.get_weights()
        np.savez(_npz_path(filepath), **{f"w{i:03d}": w for i, w in enumerate(ws)})

    def load_weights(self, filepath):
        with np.load(_npz_path(filepath)) as z:
            keys = sorted(z.files)
            self.set_weights([z[k] for k in keys])


def _npz_path(p):
    p = str(p)
    return p if p.endswith(".npz") else p + ".npz"


class BS:
    """Define the BS DNN class -- drop-in for ``BS_brain.BS`` (BS_brain.py:90-239).

    Extra keyword arguments (all optional; defaults reproduce the reference):
      stages      number of GNN stages (reference wiring: 3, :147-166)
      per_slot    one weight set per node slot as the reference instantiates them (:121-200);
                  False shares one set over all nodes (what its comment at :120 intends)
      max_batch   capacity of the device workspace (grown on demand)
      data_parallel  all-reduce gradients over torch.distributed's default group
      seed        seed of the glorot_uniform initialisation
      dtype       "f32" (reference arithmetic; default) or "bf16": bf16 contraction operands with fp32 accumulation on
                  the tensor cores for forward AND backward (csrc/tc_train.cu; BASELINE configs[2]); fp32 master
                  weights, gradients and Adam.  Shared weights, N <= 32, <= 3 stages, binary adjacency only.
    """

    def __init__(self, num_d2d, input_node_info, input_edge_info, num_d2d_feedback, num_d2d_neighbor, num_ch,
                 stages=3, per_slot=True, hidden=(80, 40, 20), max_batch=1024, data_parallel=None, seed=None, dtype="f32"):
        self.num_D2D = int(num_d2d)
        self.num_Neighbor = int(num_d2d_neighbor)
        self.num_CH = int(num_ch)
        self.num_Feedback = int(num_d2d_feedback)
        self.input_node_Info = input_node_info
        self.input_edge_Info = input_edge_info
        self.num_One_Node_Input = ((input_node_info - 1) * self.num_CH + 1) * self.num_Neighbor      # :101
        self.num_One_Edge_Input = input_edge_info * self.num_CH                                      # :102
        self.num_One_D2D_Input = self.num_One_Node_Input + self.num_One_Edge_Input                   # :103
        self.num_D2D_Input = num_d2d * self.num_One_D2D_Input + self.num_D2D ** 2                    # :104
        self.stages = int(stages)
        self.per_slot = bool(per_slot)
        self.hidden = tuple(int(h) for h in hidden)
        if str(dtype) not in ("f32", "fp32", "float32", "bf16", "bfloat16"):
            raise ValueError(f"dtype must be 'f32' or 'bf16', got {dtype!r}")
        self.dtype = "bf16" if str(dtype) in ("bf16", "bfloat16") else "f32"
        if len(self.hidden) != 3:
            raise ValueError("the decision MLP has three hidden layers (BS_brain.py:176-178)")
        self._lib = _lib.load()
        self._dev = _device()
        self._handle = None
        self._max_batch = 0
        self._pin = {}
        self._loss_keys = [f"D{k + 1}_Decide_Output_loss" for k in range(int(num_d2d))]
        if data_parallel is None:
            data_parallel = torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1
        self.data_parallel = bool(data_parallel)
        self._comm = None
        self._create(max(int(max_batch), 1))
        self._init_weights(seed)
        if self.data_parallel and os.environ.get("V2V_DP_BACKEND", "peer") == "peer":
            self.enable_peer_allreduce()
        self.model = self._create_model(0)
        self.target_model = self._create_model(1)

    # ------------------------------------------------------------------ engine lifetime
    def _create(self, max_batch, keep_state=None):
        cfg = BrainConfig()
        cfg.num_d2d, cfg.node_dim, cfg.edge_dim = self.num_D2D, self.num_One_Node_Input, self.num_One_Edge_Input
        cfg.feedback, cfg.num_ch, cfg.stages, cfg.per_slot = self.num_Feedback, self.num_CH, self.stages, int(self.per_slot)
        cfg.hidden[0], cfg.hidden[1], cfg.hidden[2] = self.hidden
        cfg.max_batch, cfg.dtype = max_batch, (_lib.V2V_BF16 if self.dtype == "bf16" else _lib.V2V_F32)
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = 1e-3, 0.5, 0.999, 1e-7           # :212, K.epsilon()
        h = C.c_void_p()
        _lib.check(self._lib.v2v_brain_create(C.byref(cfg), C.byref(h)), ValueError)
        self._handle = h
        self._max_batch = max_batch
        self.param_count = int(self._lib.v2v_brain_param_count(h))
        self._views = [torch.as_tensor(_DevBuf(self._lib.v2v_brain_param_ptr(h, w), self.param_count), device=self._dev)
                       for w in range(5)]
        if getattr(self, "_fused", 1) != 1:
            _lib.check(self._lib.v2v_brain_set_fused(h, int(self._fused)))
        if getattr(self, "_tc_mode", None) is not None:
            _lib.check(self._lib.v2v_brain_set_tensor_core(h, int(self._tc_mode)))
        if keep_state is not None:
            for w, t in enumerate(keep_state["bufs"]):
                self._views[w].copy_(t)
            _lib.check(self._lib.v2v_brain_set_iterations(h, keep_state["t"]))

    def _ensure_capacity(self, B):
        if B <= self._max_batch:
            return
        state = {"bufs": [v.clone() for v in self._views], "t": self._lib.v2v_brain_get_iterations(self._handle)}
        torch.cuda.synchronize()
        self._lib.v2v_brain_destroy(self._handle)
        self._pin = {}
        self._create(int(2 ** math.ceil(math.log2(B))), keep_state=state)

    def enable_peer_allreduce(self, group=None):
        """Replace the NCCL gradient all-reduce + Adam kernel by ONE kernel that exchanges the gradient over
        NVLink peer memory (cudaIpc-mapped buffers) and applies Adam (csrc/comm.cu).  Collective: every rank
        of the group must call it.  The handles travel through torch.distributed (plumbing)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        comm = C.c_void_p()
        _lib.check(self._lib.v2v_comm_create(self.param_count + 32, world, rank, C.byref(comm)))
        nb = self._lib.v2v_comm_ipc_handle_bytes()
        mine = C.create_string_buffer(nb)
        _lib.check(self._lib.v2v_comm_get_ipc_handle(comm, mine))
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(mine.raw), group=group)
        blob = C.create_string_buffer(b"".join(gathered), nb * world)
        _lib.check(self._lib.v2v_comm_open_peers(comm, blob))
        dist.barrier(group)
        self._comm = comm

    def __del__(self):
        try:
            if getattr(self, "_comm", None) is not None:
                torch.cuda.synchronize()
                self._lib.v2v_comm_destroy(self._comm)
                self._comm = None
            if self._handle is not None:
                self._lib.v2v_brain_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _create_model(self, which):
        return _ModelHandle(self, which)

    # ------------------------------------------------------------------ parameters
    def layer_shapes(self):
        """(K, n_out) of each stacked layer weight, GNN stages first, then the decision MLP."""
        F, Dn, De = self.num_Feedback, self.num_One_Node_Input, self.num_One_Edge_Input
        shapes = [((Dn if s == 0 else F + Dn) + De + F, F) for s in range(self.stages)]
        k = Dn + 2 * F
        for h in self.hidden:
            shapes.append((k, h))
            k = h
        shapes.append((k, self.num_CH))
        return shapes

    @property
    def groups(self):
        return self.num_D2D if self.per_slot else 1

    def _init_weights(self, seed):
        """glorot_uniform kernels / zero biases per layer object, each GNN weight with its own fan-in
        (BS_brain.py:26-41, Dense defaults).  Online and target nets are initialised independently,
        as two ``_create_model`` calls are (:105-106)."""
        rng = np.random.default_rng(seed)
        F, De = self.num_Feedback, self.num_One_Edge_Input
        for which in (0, 1):
            chunks = []
            for li, (K, O) in enumerate(self.layer_shapes()):
                W = np.empty((self.groups, K, O), np.float32)
                for g in range(self.groups):
                    if li < self.stages:
                        off = 0
                        for d in (K - De - F, De, F):
                            lim = math.sqrt(6.0 / (d + O))
                            W[g, off:off + d] = rng.uniform(-lim, lim, (d, O))
                            off += d
                    else:
                        lim = math.sqrt(6.0 / (K + O))
                        W[g] = rng.uniform(-lim, lim, (K, O))
                chunks += [W.ravel(), np.zeros(self.groups * O, np.float32)]
            self.set_flat_params(np.concatenate(chunks), which)

    def get_flat_params(self, which=0):
        """Flat fp32 copy: per layer W[G,K,O] then bias[G,O].  which: 0 online, 1 target, 2 grads, 3 m, 4 v."""
        return self._views[which].detach().cpu().numpy().copy()

    def set_flat_params(self, flat, which=0):
        flat = np.ascontiguousarray(flat, dtype=np.float32).ravel()
        if flat.size != self.param_count:
            raise ValueError(f"expected {self.param_count} parameters, got {flat.size}")
        self._views[which].copy_(torch.from_numpy(flat))

    def _get_weight_list(self, which):
        """Keras ``get_weights`` order: layer by layer, node slot by node slot, [W1,W2,W3,bias] for a
        GNNLayer (:26-41) and [kernel,bias] for a Dense."""
        flat = self.get_flat_params(which)
        F, De, G = self.num_Feedback, self.num_One_Edge_Input, self.groups
        out, o = [], 0
        for li, (K, O) in enumerate(self.layer_shapes()):
            W = flat[o:o + G * K * O].reshape(G, K, O); o += G * K * O
            b = flat[o:o + G * O].reshape(G, O); o += G * O
            for g in range(G):
                if li < self.stages:
                    da = K - De - F
                    out += [W[g, :da].copy(), W[g, da:da + De].copy(), W[g, da + De:].copy(), b[g].copy()]
                else:
                    out += [W[g].copy(), b[g].copy()]
        return out

    def _set_weight_list(self, which, weights):
        F, De, G = self.num_Feedback, self.num_One_Edge_Input, self.groups
        weights = [np.asarray(w, dtype=np.float32) for w in weights]
        per_layer = [4 if li < self.stages else 2 for li in range(len(self.layer_shapes()))]
        if len(weights) != G * sum(per_layer):
            raise ValueError(f"You called `set_weights(weights)` with a weight list of length {len(weights)}, but the "
                             f"model was expecting {G * sum(per_layer)} weights.")
        chunks, i = [], 0
        for li, (K, O) in enumerate(self.layer_shapes()):
            W = np.empty((G, K, O), np.float32)
            b = np.empty((G, O), np.float32)
            for g in range(G):
                if li < self.stages:
                    stacked = np.concatenate(weights[i:i + 3], axis=0)
                    bias = weights[i + 3]
                    i += 4
                else:
                    stacked, bias = weights[i], weights[i + 1]
                    i += 2
                if stacked.shape != (K, O) or bias.shape != (O,):
                    raise ValueError(f"Layer weight shape {(K, O)} not compatible with provided weight shape "
                                     f"{stacked.shape}")
                W[g], b[g] = stacked, bias
            chunks += [W.ravel(), b.ravel()]
        self.set_flat_params(np.concatenate(chunks), which)

    def set_fused(self, enable):
        """Shared-weight brains run forward/backward as one fused kernel; ``False`` forces the
        layer-by-layer kernels (always used for per-slot weights, weighted adjacency, N > 32)."""
        mode = int(bool(enable))
        _lib.check(self._lib.v2v_brain_set_fused(self._handle, mode))
        self._fused = mode

    def set_tensor_core(self, mode):
        """predict on the tcgen05 tensor cores (3xTF32, fp32-grade): 0 never, 1 automatic (large batches), 2 always."""
        _lib.check(self._lib.v2v_brain_set_tensor_core(self._handle, int(mode)), ValueError)
        self._tc_mode = int(mode)

    def tensor_core_info(self):
        info = (C.c_int32 * 4)()
        _lib.check(self._lib.v2v_brain_tensor_core_info(self._handle, info))
        return dict(zip(("capable", "mode", "graphs_per_tile", "smem_bytes"), list(info)))

    def fused_info(self, B, train=True):
        info = (C.c_int * 8)()
        _lib.check(self._lib.v2v_brain_fused_info(self._handle, int(B), int(bool(train)), info))
        keys = ("capable", "graphs_per_tile", "arena_rows", "smem_bytes", "phases", "wgrad_blocks", "bias_slots", "table")
        return dict(zip(keys, list(info)))

    @property
    def iterations(self):
        return int(self._lib.v2v_brain_get_iterations(self._handle))

    # ------------------------------------------------------------------ input adaptation
    def _pinned(self, name, shape):
        t = self._pin.get(name)
        n = int(np.prod(shape))
        if t is None or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=torch.float32, pin_memory=True)
            self._pin[name] = t
        return t[:n].view(*shape)

    def _pack_inputs(self, data):
        """dict of per-slot arrays (BS_brain.py:495-504) -> pinned node/edge/(neighbor)/adj buffers."""
        N, Dn, De, F = self.num_D2D, self.num_One_Node_Input, self.num_One_Edge_Input, self.num_Feedback
        if not isinstance(data, dict):
            raise ValueError("inputs must be a dict keyed by the model's input names")
        if "Node_Input" in data:                       # packed extension: (B,N,Dn), (B,N,De)
            node_src = np.asarray(data["Node_Input"])
            edge_src = np.asarray(data["Edge_Input"])
            B = node_src.shape[0]
            if node_src.shape != (B, N, Dn) or edge_src.shape != (B, N, De):
                raise ValueError(f"Node_Input/Edge_Input must be (B,{N},{Dn}) / (B,{N},{De})")
            node = self._pinned("node", (B, N, Dn)); node.numpy()[...] = node_src
            edge = self._pinned("edge", (B, N, De)); edge.numpy()[...] = edge_src
            neigh_src = data.get("Neighbor_Input")
            neigh = None
            if neigh_src is not None and np.any(neigh_src):
                neigh = self._pinned("neigh", (B, N, F)); neigh.numpy()[...] = np.asarray(neigh_src)
        else:
            missing = [f"D{k + 1}_{kind}_Input" for k in range(N) for kind in ("Node", "Edge")
                       if f"D{k + 1}_{kind}_Input" not in data]
            if missing or "Adjacency_Matrix" not in data:
                raise ValueError(f"No data provided for \"{(missing + ['Adjacency_Matrix'])[0]}\". Need data for each key "
                                 f"in the model's inputs")
            first = np.asarray(data["D1_Node_Input"])
            B = first.shape[0]
            node = self._pinned("node", (B, N, Dn)); edge = self._pinned("edge", (B, N, De))
            nv, ev = node.numpy(), edge.numpy()
            neigh = None
            for k in range(N):
                a = np.asarray(data[f"D{k + 1}_Node_Input"]); e = np.asarray(data[f"D{k + 1}_Edge_Input"])
                if a.shape != (B, Dn):
                    raise ValueError(f"Error when checking input: expected D{k + 1}_Node_Input to have shape ({Dn},) "
                                     f"but got array with shape {a.shape[1:]}")
                if e.shape != (B, De):
                    raise ValueError(f"Error when checking input: expected D{k + 1}_Edge_Input to have shape ({De},) "
                                     f"but got array with shape {e.shape[1:]}")
                nv[:, k, :] = a
                ev[:, k, :] = e
            nb = [data.get(f"D{k + 1}_Neighbor_Input") for k in range(N)]
            if any(x is not None and np.any(x) for x in nb):     # the reference always feeds zeros (:478, :589)
                neigh = self._pinned("neigh", (B, N, F))
                for k in range(N):
                    neigh.numpy()[:, k, :] = 0 if nb[k] is None else np.asarray(nb[k])
        A = np.asarray(data["Adjacency_Matrix"])
        if A.ndim != 3 or A.shape[0] != B:
            raise ValueError(f"Adjacency_Matrix must be (B, N*F, N*F) or (B, N, N) with B={B}, got {A.shape}")
        if A.shape[1:] == (N, N):
            adj_src = A
        elif A.shape[1:] == (N * F, N * F):
            adj_src = A[:, ::F, ::F]                    # exact inverse of kron(Adj, I_F) (:492-493)
        else:
            raise ValueError(f"Error when checking input: expected Adjacency_Matrix to have shape ({N * F}, {N * F}) "
                             f"but got array with shape {A.shape[1:]}")
        adj = self._pinned("adj", (B, N, N)); adj.numpy()[...] = adj_src
        return B, node, edge, neigh, adj

    def _pack_labels(self, labels, B):
        N, CH = self.num_D2D, self.num_CH
        y = self._pinned("y", (B, N, CH))
        if isinstance(labels, dict):
            if "Decide_Output" in labels:
                y.numpy()[...] = np.asarray(labels["Decide_Output"])
                return y
            for k in range(N):
                key = f"D{k + 1}_Decide_Output"
                if key not in labels:
                    raise ValueError(f"No data provided for \"{key}\". Need data for each key in the model's outputs")
                arr = np.asarray(labels[key])
                if arr.shape != (B, CH):
                    raise ValueError(f"Error when checking target: expected {key} to have shape ({CH},) but got array "
                                     f"with shape {arr.shape[1:]}")
                y.numpy()[:, k, :] = arr
        else:
            labels = list(labels)
            if len(labels) != N:
                raise ValueError(f"expected {N} target arrays, got {len(labels)}")
            for k in range(N):
                y.numpy()[:, k, :] = np.asarray(labels[k])
        return y

    # ------------------------------------------------------------------ zero-copy input description
    @staticmethod
    def _view(arr, rows, cols, dst_off, dst_row_stride, keep, row_stride=None, col_stride=None):
        """One ``v2v_host_view`` over a 2-D (rows, cols) numpy window; non-fp32/fp64 data is converted first."""
        if arr.dtype != np.float32 and arr.dtype != np.float64:
            arr = arr.astype(np.float32)
        it = arr.itemsize
        if row_stride is None:
            rs, cs = arr.strides[0], arr.strides[1]
            if rs % it or cs % it:
                arr = np.ascontiguousarray(arr)
                rs, cs = arr.strides
            row_stride, col_stride = rs // it, cs // it
        keep.append(arr)                                   # the C side reads the caller's memory during the call
        return _lib.HostView(arr.__array_interface__["data"][0], _lib.V2V_F32 if it == 4 else _lib.V2V_F64, rows, cols,
                             row_stride, col_stride, dst_off, dst_row_stride)

    def _slot_views(self, data, kind, width, keep, B=None, packed_key=None):
        """Views of one input tensor [B][N][width]: the packed extension array, or the reference's per-slot arrays."""
        N = self.num_D2D
        if packed_key is not None and packed_key in data:
            a = np.asarray(data[packed_key])
            if a.ndim != 3 or a.shape[1:] != (N, width) or (B is not None and a.shape[0] != B):
                raise ValueError(f"{packed_key} must be (B,{N},{width}), got {a.shape}")
            B = a.shape[0]
            if not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)
            return B, [self._view(a.reshape(B, N * width), B, N * width, 0, N * width, keep)]
        views = []
        for k in range(N):
            key = kind.format(k + 1)
            if key not in data:
                raise ValueError(f"No data provided for \"{key}\". Need data for each key in the model's "
                                 f"{'outputs' if 'Output' in key else 'inputs'}")
            a = np.asarray(data[key])
            if B is None:
                B = a.shape[0]
            if a.shape != (B, width):
                what = "target" if "Output" in key else "input"
                raise ValueError(f"Error when checking {what}: expected {key} to have shape ({width},) but got array "
                                 f"with shape {a.shape[1:]}")
            views.append(self._view(a, B, width, k * width, N * width, keep))
        return B, views

    def _input_views(self, data, keep):
        """dict of numpy inputs (BS_brain.py:495-504, or the packed extension keys) -> ctypes view arrays."""
        N, Dn, De, F = self.num_D2D, self.num_One_Node_Input, self.num_One_Edge_Input, self.num_Feedback
        if not isinstance(data, dict):
            raise ValueError("inputs must be a dict keyed by the model's input names")
        packed = "Node_Input" in data
        if not packed and "Adjacency_Matrix" not in data and all(f"D{k + 1}_Node_Input" in data for k in range(N)):
            raise ValueError("No data provided for \"Adjacency_Matrix\". Need data for each key in the model's inputs")
        B, node = self._slot_views(data, "D{}_Node_Input", Dn, keep, packed_key="Node_Input" if packed else None)
        _, edge = self._slot_views(data, "D{}_Edge_Input", De, keep, B=B, packed_key="Edge_Input" if packed else None)
        # neighbour inputs: the reference always feeds zeros (:478, :589); the C side notices (non-zero scan while
        # gathering) and neither ships them nor runs that contraction
        neigh = []
        if packed:
            if data.get("Neighbor_Input") is not None:
                _, neigh = self._slot_views(data, "", F, keep, B=B, packed_key="Neighbor_Input")
        else:
            nbs = [data.get(f"D{k + 1}_Neighbor_Input") for k in range(N)]
            if all(x is not None for x in nbs):
                _, neigh = self._slot_views(data, "D{}_Neighbor_Input", F, keep, B=B)
            elif any(x is not None for x in nbs):
                full = {f"D{k + 1}_Neighbor_Input": (np.zeros((B, F), np.float32) if nbs[k] is None else nbs[k])
                        for k in range(N)}
                _, neigh = self._slot_views(full, "D{}_Neighbor_Input", F, keep, B=B)
        if "Adjacency_Matrix" not in data:
            raise ValueError("No data provided for \"Adjacency_Matrix\". Need data for each key in the model's inputs")
        A = np.asarray(data["Adjacency_Matrix"])
        if A.ndim != 3 or A.shape[0] != B:
            raise ValueError(f"Adjacency_Matrix must be (B, N*F, N*F) or (B, N, N) with B={B}, got {A.shape}")
        if A.dtype != np.float32 and A.dtype != np.float64:
            A = A.astype(np.float32)
        if not A.flags.c_contiguous:
            A = np.ascontiguousarray(A)
        if A.shape[1:] == (N, N):
            adj = [self._view(A, B * N, N, 0, N, keep, row_stride=N, col_stride=1)]
        elif A.shape[1:] == (N * F, N * F):             # kron(Adj, I_F) (:492-493): sample every F-th row and column
            adj = [self._view(A, B * N, N, 0, N, keep, row_stride=F * N * F, col_stride=F)]
        else:
            raise ValueError(f"Error when checking input: expected Adjacency_Matrix to have shape ({N * F}, {N * F}) "
                             f"but got array with shape {A.shape[1:]}")
        return B, node, edge, neigh, adj

    def _label_views(self, labels, B, keep):
        N, CH = self.num_D2D, self.num_CH
        if not isinstance(labels, dict):
            labels = list(labels)
            if len(labels) != N:
                raise ValueError(f"expected {N} target arrays, got {len(labels)}")
            labels = {f"D{k + 1}_Decide_Output": labels[k] for k in range(N)}
        packed = "Decide_Output" in labels
        _, y = self._slot_views(labels, "D{}_Decide_Output", CH, keep, B=B, packed_key="Decide_Output" if packed else None)
        return y

    @staticmethod
    def _varr(views):
        return (_lib.HostView * max(len(views), 1))(*views), len(views)

    # ------------------------------------------------------------------ the reference's methods
    def predict(self, data_test, target=False):
        """BS.predict (BS_brain.py:225-231): list of ``num_D2D`` writable arrays (B, num_CH)."""
        keep = []
        B, node, edge, neigh, adj = self._input_views(data_test, keep)
        self._ensure_capacity(B)
        q = np.empty((B, self.num_D2D, self.num_CH), np.float32)
        (na, nn), (ea, en), (ga, gn), (aa, an) = (self._varr(v) for v in (node, edge, neigh, adj))
        _lib.check(self._lib.v2v_brain_predict_views(self._handle, na, nn, ea, en, ga, gn, aa, an, B, int(bool(target)),
                                                     q.ctypes.data, _lib.current_stream()), ValueError)
        return [q[:, k, :].copy() for k in range(self.num_D2D)]

    def predict_one_step(self, data_test, target=False):
        """BS.predict_one_step (BS_brain.py:233-235)."""
        return self.predict(data_test, target=target)

    def update_target_model(self):
        """BS.update_target_model (BS_brain.py:237-239)."""
        _lib.check(self._lib.v2v_brain_update_target(self._handle, _lib.current_stream()))

    def train_dnn(self, data_train, labels, batch_size):
        """BS.train_dnn (BS_brain.py:218-223): ``model.fit(..., batch_size, epochs=1, verbose=0)``."""
        return self.model.fit(data_train, labels, batch_size=batch_size, epochs=1, verbose=0)

    def _fit(self, x, y, batch_size, epochs, shuffle):
        bs_arg = None if batch_size is None else int(batch_size)
        if bs_arg is not None and bs_arg < 1:
            raise ValueError("batch_size must be >= 1")
        if not self.data_parallel or self._comm is not None:
            # one optimiser step per epoch on all rows (what the reference always does: :566-567, :728): the C side
            # gathers the caller's arrays itself (csrc/host_stage.cu), nothing is repacked in Python
            keep = []
            B, node, edge, neigh, adj = self._input_views(x, keep)
            if bs_arg is None or B <= bs_arg:
                yv = self._label_views(y, B, keep)
                self._ensure_capacity(B)
                (na, nn), (ea, en), (ga, gn), (aa, an), (ya, yn) = (self._varr(v) for v in (node, edge, neigh, adj, yv))
                N = self.num_D2D
                hl = np.empty(N, np.float32)
                hist = History()
                keys = self._loss_keys
                history = {"loss": [], **{k: [] for k in keys}} if int(epochs) != 1 else None
                for ep in range(int(epochs)):
                    if self._comm is not None:       # data parallel: this rank's rows, fused NVLink exchange + Adam
                        rc = self._lib.v2v_brain_train_views_dp(self._handle, self._comm, na, nn, ea, en, ga, gn, aa, an,
                                                                ya, yn, B, hl.ctypes.data, _lib.current_stream())
                    else:
                        rc = self._lib.v2v_brain_train_views(self._handle, na, nn, ea, en, ga, gn, aa, an, ya, yn, B,
                                                             hl.ctypes.data, _lib.current_stream())
                    _lib.check(rc, ValueError)
                    per_head = hl.tolist()                         # float32 values as Python floats (exact)
                    hist.epoch.append(ep)
                    if history is None:                             # the reference's only case: epochs = 1 (:220-221)
                        history = {"loss": [float(sum(per_head))]}
                        history.update(zip(keys, ([v] for v in per_head)))
                    else:
                        history["loss"].append(float(sum(per_head)))
                        for k in range(N):
                            history[keys[k]].append(per_head[k])
                hist.history = history if history is not None else {"loss": [], **{k: [] for k in keys}}
                return hist
        B, node, edge, neigh, adj = self._pack_inputs(x)
        ylab = self._pack_labels(y, B)
        bs = B if batch_size is None else int(batch_size)
        hist = History()
        N = self.num_D2D
        keys = [f"D{k + 1}_Decide_Output_loss" for k in range(N)]
        hist.history = {"loss": [], **{k: [] for k in keys}}
        for ep in range(int(epochs)):
            if B <= bs:
                per_head = self._train_rows(node, edge, neigh, adj, ylab, B)
            else:                                        # Keras mini-batching over a shuffled index array
                idx = np.random.permutation(B) if shuffle else np.arange(B)
                acc = np.zeros(N, np.float64)
                for s in range(0, B, bs):
                    sel = torch.from_numpy(np.sort(idx[s:s + bs]) if not shuffle else idx[s:s + bs])
                    nb = len(sel)
                    sub = [self._pinned("mb_" + nm, (nb,) + tuple(t.shape[1:])) for nm, t in
                           (("node", node), ("edge", edge), ("adj", adj), ("y", ylab))]
                    for dst, src in zip(sub, (node, edge, adj, ylab)):
                        torch.index_select(src, 0, sel, out=dst)
                    nsub = None
                    if neigh is not None:
                        nsub = self._pinned("mb_neigh", (nb,) + tuple(neigh.shape[1:]))
                        torch.index_select(neigh, 0, sel, out=nsub)
                    acc += self._train_rows(sub[0], sub[1], nsub, sub[2], sub[3], nb) * nb
                per_head = acc / B
            hist.epoch.append(ep)
            hist.history["loss"].append(float(np.sum(per_head)))
            for k in range(N):
                hist.history[keys[k]].append(float(per_head[k]))
        return hist

    def _train_rows(self, node, edge, neigh, adj, y, B):
        """One optimiser step on exactly B rows from pinned host buffers; returns per-head losses."""
        self._ensure_capacity(B)
        hl = self._pinned("head_loss", (self.num_D2D,))
        if not self.data_parallel:
            _lib.check(self._lib.v2v_brain_train_host(self._handle, ptr(node), ptr(edge), ptr(neigh), ptr(adj), ptr(y), B,
                                                      ptr(hl), _lib.current_stream()), ValueError)
            return hl.numpy().astype(np.float64)
        # data parallel: local fwd+bwd, one all-reduce of [grads | head losses], identical Adam everywhere
        dev = self._dev
        nd, ed, ad, yd = (t.to(dev, non_blocking=True) for t in (node, edge, adj, y))
        ngd = None if neigh is None else neigh.to(dev, non_blocking=True)
        from .layers import pack_adjacency
        in_mask, out_mask, binary = pack_adjacency(ad)
        losses = self.train_step_device(nd, ed, in_mask if binary else None, out_mask if binary else None,
                                        None if binary else ad, yd, neighbor=ngd)
        return losses.cpu().numpy().astype(np.float64)

    # ------------------------------------------------------------------ device-resident fast path
    def _check_dev(self, what, t, shape, dtype, optional=False):
        """The device entry points take raw pointers: a wrong dtype / stride / shape would read out of bounds."""
        if t is None:
            if optional:
                return
            raise ValueError(f"{what}: tensor required")
        if not t.is_cuda or t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise ValueError(f"{what}: expected a contiguous CUDA {dtype} tensor of shape {tuple(shape)}, got "
                             f"{t.dtype} {tuple(t.shape)} (cuda={t.is_cuda}, contiguous={t.is_contiguous()})")

    def _check_graph_inputs(self, who, node, edge, in_mask, out_mask, adj, neighbor):
        B, N = int(node.shape[0]), self.num_D2D
        W = (N + 31) // 32
        f32, i32 = torch.float32, torch.int32
        self._check_dev(f"{who}: node", node, (B, N, self.num_One_Node_Input), f32)
        self._check_dev(f"{who}: edge", edge, (B, N, self.num_One_Edge_Input), f32)
        self._check_dev(f"{who}: neighbor", neighbor, (B, N, self.num_Feedback), f32, optional=True)
        self._check_dev(f"{who}: adj", adj, (B, N, N), f32, optional=True)
        for nm, m in (("in_mask", in_mask), ("out_mask", out_mask)):
            if m is not None and (not m.is_cuda or m.dtype != i32 or not m.is_contiguous() or m.numel() != B * N * W):
                raise ValueError(f"{who}: {nm} must be a contiguous CUDA int32 tensor of {B * N * W} words, got {m.dtype} "
                                 f"{tuple(m.shape)}")
        return B

    def forward_device(self, node, edge, in_mask=None, adj=None, target=False, neighbor=None, out=None):
        """Device tensors in, Q [B,N,CH] device tensor out (no host round trip)."""
        B = self._check_graph_inputs("forward_device", node, edge, in_mask, None, adj, neighbor)
        self._ensure_capacity(B)
        if out is None:
            out = torch.empty((B, self.num_D2D, self.num_CH), dtype=torch.float32, device=node.device)
        else:
            self._check_dev("forward_device: out", out, (B, self.num_D2D, self.num_CH), torch.float32)
        _lib.check(self._lib.v2v_brain_forward(self._handle, ptr(node), ptr(edge), ptr(neighbor), ptr(in_mask), ptr(adj), B,
                                               int(bool(target)), ptr(out), _lib.current_stream()), ValueError)
        return out

    def train_step_device(self, node, edge, in_mask, out_mask, adj, y, neighbor=None, head_loss=None):
        """One fwd+bwd(+all-reduce)+Adam step on device tensors; returns per-head loss tensor [N] (device)."""
        B = self._check_graph_inputs("train_step_device", node, edge, in_mask, out_mask, adj, neighbor)
        self._ensure_capacity(B)
        N = self.num_D2D
        self._check_dev("train_step_device: y", y, (B, N, self.num_CH), torch.float32)
        st = _lib.current_stream()
        if head_loss is None:
            head_loss = torch.empty(N, dtype=torch.float32, device=node.device)
        else:
            self._check_dev("train_step_device: head_loss", head_loss, (N,), torch.float32)
        if not self.data_parallel:
            _lib.check(self._lib.v2v_brain_train_step(self._handle, ptr(node), ptr(edge), ptr(neighbor), ptr(in_mask),
                                                      ptr(out_mask), ptr(adj), ptr(y), B, ptr(head_loss), st), ValueError)
            return head_loss
        if self._comm is not None:              # fused reduce + NVLink exchange + Adam (one kernel)
            _lib.check(self._lib.v2v_brain_train_step_dp(self._handle, self._comm, ptr(node), ptr(edge), ptr(neighbor),
                                                         ptr(in_mask), ptr(out_mask), ptr(adj), ptr(y), B, ptr(head_loss),
                                                         st), ValueError)
            return head_loss
        import torch.distributed as dist
        _lib.check(self._lib.v2v_brain_forward_backward(self._handle, ptr(node), ptr(edge), ptr(neighbor), ptr(in_mask),
                                                        ptr(out_mask), ptr(adj), ptr(y), B, ptr(head_loss), st), ValueError)
        # the one collective of the step; AVG leaves the global-batch mean gradient in the grad buffer (as the peer path does)
        dist.all_reduce(self._views[2], op=dist.ReduceOp.AVG)
        dist.all_reduce(head_loss, op=dist.ReduceOp.AVG)          # per-head losses of the global batch, like the peer path
        _lib.check(self._lib.v2v_brain_apply_adam(self._handle, 1.0, st))
        return head_loss
