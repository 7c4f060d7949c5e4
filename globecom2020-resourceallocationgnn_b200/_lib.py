"""ctypes binding of the C-ABI in include/v2v_gnn.h.

There is no CPU or eager-PyTorch fallback: if the shared library cannot be
loaded (and cannot be built because nvcc is absent) importing any compute entry
point raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
c_i32_p = C.POINTER(C.c_int32)
c_void_p = C.c_void_p


class BrainConfig(C.Structure):
    """Mirror of ``v2v_brain_config`` (include/v2v_gnn.h)."""
    _fields_ = [
        ("num_d2d", C.c_int), ("node_dim", C.c_int), ("edge_dim", C.c_int), ("feedback", C.c_int),
        ("num_ch", C.c_int), ("stages", C.c_int), ("per_slot", C.c_int), ("hidden", C.c_int * 3),
        ("max_batch", C.c_int), ("dtype", C.c_int),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
    ]


class HostView(C.Structure):
    """Mirror of ``v2v_host_view`` (include/v2v_gnn.h): a strided window of caller memory and where it lands."""
    _fields_ = [
        ("ptr", C.c_void_p), ("dtype", C.c_int), ("rows", C.c_long), ("cols", C.c_long),
        ("row_stride", C.c_long), ("col_stride", C.c_long), ("dst_off", C.c_long), ("dst_row_stride", C.c_long),
    ]


V2V_F32, V2V_BF16, V2V_F64 = 0, 1, 2
c_view_p = C.POINTER(HostView)

# name -> (restype, argtypes).  Every symbol include/v2v_gnn.h declares is listed here;
# tests/test_capi_symbols.py checks the two stay in sync.
SIGNATURES = {
    "v2v_last_error": (C.c_char_p, []),
    "v2v_version": (C.c_int, []),
    "v2v_device_sm_count": (C.c_int, []),
    "v2v_launch_count": (C.c_long, []),
    "v2v_adj_pack_masks": (C.c_int, [c_void_p, C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "v2v_agg_mask": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_agg_mask_ex": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint,
                                  c_void_p]),
    "v2v_agg_dense": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_dense_fwd": (C.c_int, [C.c_int, C.POINTER(c_void_p), C.POINTER(C.c_int), c_void_p, C.c_int, c_void_p,
                                c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_dense_bwd_data": (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p, C.c_int,
                                     C.c_int, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_dense_bwd_weight": (C.c_int, [C.c_int, C.POINTER(c_void_p), C.POINTER(C.c_int), c_void_p, c_void_p,
                                       c_void_p, C.c_int, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_huber_loss_grad": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_float, c_void_p]),
    "v2v_dqn_select_actions": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_dqn_replay_write": (C.c_int, [c_void_p] * 19 + [C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_td_target": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_float, c_void_p, C.c_int, C.c_int,
                                C.c_int, c_void_p]),
    "v2v_adam_step": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_long, C.c_int, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, c_void_p]),
    "v2v_brain_create": (C.c_int, [C.POINTER(BrainConfig), C.POINTER(c_void_p)]),
    "v2v_brain_destroy": (None, [c_void_p]),
    "v2v_brain_param_count": (C.c_long, [c_void_p]),
    "v2v_brain_get_params": (C.c_int, [c_void_p, C.c_int, c_void_p, c_void_p]),
    "v2v_brain_set_params": (C.c_int, [c_void_p, C.c_int, c_void_p, c_void_p]),
    "v2v_brain_param_ptr": (c_void_p, [c_void_p, C.c_int]),
    "v2v_brain_set_fused": (C.c_int, [c_void_p, C.c_int]),
    "v2v_brain_fused_info": (C.c_int, [c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "v2v_fused_set_trace": (C.c_int, [c_void_p]),
    "v2v_fused_set_mma": (C.c_int, [C.c_int]),
    "v2v_fused_get_mma": (C.c_int, []),
    "v2v_fused_plan": (C.c_int, [C.POINTER(BrainConfig), C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "v2v_brain_update_target": (C.c_int, [c_void_p, c_void_p]),
    "v2v_brain_get_iterations": (C.c_int, [c_void_p]),
    "v2v_brain_set_iterations": (C.c_int, [c_void_p, C.c_int]),
    "v2v_brain_forward": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, c_void_p,
                                    c_void_p]),
    "v2v_brain_forward_backward": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             C.c_int, c_void_p, c_void_p]),
    "v2v_brain_apply_adam": (C.c_int, [c_void_p, C.c_float, c_void_p]),
    "v2v_comm_create": (C.c_int, [C.c_long, C.c_int, C.c_int, C.POINTER(c_void_p)]),
    "v2v_comm_destroy": (None, [c_void_p]),
    "v2v_comm_ipc_handle_bytes": (C.c_int, []),
    "v2v_comm_get_ipc_handle": (C.c_int, [c_void_p, c_void_p]),
    "v2v_comm_open_peers": (C.c_int, [c_void_p, c_void_p]),
    "v2v_comm_allreduce_adam": (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_long, c_void_p, C.c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                          c_void_p]),
    "v2v_comm_allreduce_adam_ex": (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_long, C.c_long, C.c_long, c_void_p, C.c_int,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_float,
                                             C.c_float, C.c_float, C.c_float, c_void_p]),
    "v2v_comm_check": (C.c_int, [c_void_p, c_void_p]),
    "v2v_comm_poll_error": (C.c_int, [c_void_p, c_void_p]),
    "v2v_comm_poll_result": (C.c_int, [c_void_p]),
    "v2v_comm_set_trace": (C.c_int, [c_void_p, c_void_p]),
    "v2v_comm_num_chunks": (C.c_int, [c_void_p]),
    "v2v_brain_train_step_dp": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, C.c_int, c_void_p, c_void_p]),
    "v2v_brain_train_step": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       C.c_int, c_void_p, c_void_p]),
    "v2v_brain_predict_host": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, c_void_p,
                                         c_void_p]),
    "v2v_brain_train_host": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, c_void_p,
                                       c_void_p]),
    "v2v_brain_predict_views": (C.c_int, [c_void_p, c_view_p, C.c_int, c_view_p, C.c_int, c_view_p, C.c_int, c_view_p,
                                          C.c_int, C.c_int, C.c_int, c_void_p, c_void_p]),
    "v2v_brain_train_views": (C.c_int, [c_void_p, c_view_p, C.c_int, c_view_p, C.c_int, c_view_p, C.c_int, c_view_p,
                                        C.c_int, c_view_p, C.c_int, C.c_int, c_void_p, c_void_p]),
    "v2v_brain_train_views_dp": (C.c_int, [c_void_p, c_void_p, c_view_p, C.c_int, c_view_p, C.c_int, c_view_p, C.c_int,
                                           c_view_p, C.c_int, c_view_p, C.c_int, C.c_int, c_void_p, c_void_p]),
    "v2v_env_renew_channels": (C.c_int, [c_void_p] * 11 + [C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_env_pack_state": (C.c_int, [c_void_p] * 8 + [C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_env_reward": (C.c_int, [c_void_p] * 9 + [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, c_void_p]),
    "v2v_env_renew_positions": (C.c_int, [c_void_p] * 4 + [C.c_int, C.c_int, c_void_p]),
    "v2v_env_choose_destinations": (C.c_int, [c_void_p] * 3 + [C.c_int, C.c_int, c_void_p]),
    "v2v_host_stage_threads": (C.c_int, []),
    "v2v_brain_set_tensor_core": (C.c_int, [c_void_p, C.c_int]),
    "v2v_brain_tensor_core_info": (C.c_int, [c_void_p, c_i32_p]),
    "v2v_tc_plan": (C.c_int, [C.POINTER(BrainConfig), c_i32_p]),
    "v2v_tt_plan": (C.c_int, [C.POINTER(BrainConfig), c_i32_p]),
    "v2v_tt_set_trace": (C.c_int, [c_void_p]),
    "v2v_brain_tc_debug": (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, c_void_p, c_void_p, c_i32_p,
                                     c_void_p]),
    "v2v_host_gather": (C.c_int, [c_view_p, C.c_int, c_void_p, C.c_long, C.c_int, c_i32_p]),
    "v2v_host_pack_adjacency": (C.c_int, [c_view_p, C.c_int, C.c_int, c_void_p, c_void_p, c_i32_p]),
}


class V2VError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent and nvcc exists). Never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        _build.build()
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, exc=V2VError):
    if rc != 0:
        msg = load().v2v_last_error()
        raise exc(msg.decode() if msg else f"v2v error {rc}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
