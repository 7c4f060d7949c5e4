"""The C-ABI library loads on a CPU-only box and exports every symbol include/v2v_gnn.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "v2v_gnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(v2v_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("v2v_agg_mask", "v2v_dense_fwd", "v2v_dense_bwd_data", "v2v_dense_bwd_weight", "v2v_huber_loss_grad",
                 "v2v_adam_step", "v2v_td_target", "v2v_brain_create", "v2v_brain_train_host", "v2v_brain_predict_host"):
        assert must in names


def test_library_exports_every_declared_symbol(v2v):
    lib = ctypes.CDLL(v2v.lib_path())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_bindings_cover_the_header(v2v):
    from importlib import import_module
    sig = import_module("globecom2020-resourceallocationgnn_b200._lib").SIGNATURES
    assert sorted(sig) == _declared()


def test_last_error_and_argument_validation_without_gpu(v2v):
    lib = v2v.load_library()
    assert lib.v2v_version() >= 100
    # shape validation happens before any CUDA call: usable as a host-only check
    rc = lib.v2v_agg_mask(None, None, None, None, 4, 0, 16, 0, None)
    assert rc != 0 and b"bad shape" in lib.v2v_last_error()
    rc = lib.v2v_dense_fwd(0, None, None, None, 0, None, None, 1, 4, 1, 16, 0, None)
    assert rc != 0 and b"n_seg" in lib.v2v_last_error()
    rc = lib.v2v_dense_fwd(1, None, None, None, 0, None, None, 1, 4, 3, 16, 0, None)
    assert rc != 0 and b"G must be" in lib.v2v_last_error()


def test_backward_pipe_switch_is_a_host_side_setting(v2v):
    """v2v_fused_set_mma / v2v_fused_get_mma (which pipe runs the backward contractions of the shared-weight fp32 kernel):
    process-wide, validated, default = tensor cores unless V2V_FUSED_MMA says otherwise.  No device needed."""
    import os
    lib = v2v.load_library()
    default = lib.v2v_fused_get_mma()
    assert default == (int(os.environ["V2V_FUSED_MMA"]) if os.environ.get("V2V_FUSED_MMA") in ("0", "1") else 1)
    try:
        for mode in (0, 1):
            assert lib.v2v_fused_set_mma(mode) == 0 and lib.v2v_fused_get_mma() == mode
        for bad in (-1, 2, 7):
            assert lib.v2v_fused_set_mma(bad) != 0 and b"outside [0,1]" in lib.v2v_last_error()
            assert lib.v2v_fused_get_mma() == 1            # a rejected value leaves the setting alone
    finally:
        assert lib.v2v_fused_set_mma(default) == 0
