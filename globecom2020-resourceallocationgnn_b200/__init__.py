"""B200-native V2V graph-convolution engine: drop-in for the hot path of
Coolzyh/Globecom2020-ResourceAllocationGNN (``BS_brain.py:17-239``).

    import importlib
    v2v = importlib.import_module("globecom2020-resourceallocationgnn_b200")   # or: import v2v_gnn_b200 as v2v
    brain = v2v.BS(4, 3, 1, 16, 1, 4)            # same signature as BS_brain.BS (:94)

The package directory carries the reference repository's name (hyphens included), so
it is imported through importlib or the ``v2v_gnn_b200`` alias module at the repo root.
"""
from . import _lib, build                                     # noqa: F401
from ._lib import V2VError, load as load_library, lib_path    # noqa: F401
from .layers import GNNLayer, AggLayer, aggregate, pack_adjacency, adjacency_from_input, dense_forward  # noqa: F401
from .brain import BS, History                                # noqa: F401
from .dqn import Agent, BatchedAgent, ReplayRing, pack_state, get_state     # noqa: F401
from .env import BatchedEnviron                                # noqa: F401

__all__ = ["GNNLayer", "AggLayer", "BS", "History", "Agent", "BatchedAgent", "ReplayRing", "BatchedEnviron", "pack_state", "get_state", "aggregate", "pack_adjacency", "adjacency_from_input",
           "dense_forward", "load_library", "lib_path", "V2VError", "huber_loss"]


def huber_loss(y_true, y_pred):
    """``huber_loss`` of BS_brain.py:86-87 (tf.losses.huber_loss, delta 1, mean over all elements),
    evaluated by the engine's loss kernel.  Accepts numpy arrays or CUDA tensors shaped (B, CH)."""
    import torch
    from .layers import _to_dev
    from ._lib import ptr
    lib = _lib.load()
    q, was_np = _to_dev(y_pred)
    y, _ = _to_dev(y_true)
    q2 = q.reshape(q.shape[0], 1, -1).contiguous()
    y2 = y.reshape(y.shape[0], 1, -1).contiguous()
    dq = torch.empty_like(q2)
    out = torch.zeros(1, dtype=torch.float32, device=q2.device)
    _lib.check(lib.v2v_huber_loss_grad(ptr(q2), ptr(y2), ptr(dq), ptr(out), q2.shape[0], 1, q2.shape[2], 1.0,
                                       _lib.current_stream()), ValueError)
    return float(out.item()) if was_np else out[0]
