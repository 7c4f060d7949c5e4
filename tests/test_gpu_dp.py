"""Data-parallel path on GPUs: the fused reduce + NVLink exchange + Adam kernel (csrc/comm.cu)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import v2v_oracle as O

pytestmark = pytest.mark.gpu


def test_allreduce_adam_kernel_world1_matches_oracle(v2v):
    """world = 1: the kernel degenerates to partial reduction + Keras-Adam; checked against the fp64 rule."""
    lib = v2v.load_library()
    rng = np.random.default_rng(0)
    n, n_cta, n_extra = 8720, 37, 20
    comm = C.c_void_p()
    assert lib.v2v_comm_create(n + 32, 1, 0, C.byref(comm)) == 0
    partial = rng.normal(size=(n_cta, n)).astype(np.float32)
    extra = rng.normal(size=n_extra).astype(np.float32)
    p0 = rng.normal(size=n).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).cuda()
    pd_, p, m, v, g = dev(partial), dev(p0), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ex, exo = dev(extra), torch.zeros(n_extra, device="cuda")
    rp, rm, rv = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    gref = partial.astype(np.float64).sum(0)
    for t in (1, 2, 3):
        rc = lib.v2v_comm_allreduce_adam(comm, pd_.data_ptr(), n_cta, n, ex.data_ptr(), n_extra, g.data_ptr(), p.data_ptr(),
                                         m.data_ptr(), v.data_ptr(), exo.data_ptr(), t, 1e-3, 0.5, 0.999, 1e-7, None)
        assert rc == 0, lib.v2v_last_error()
        rp, rm, rv = O.keras_adam_step(rp, gref, rm, rv, t)
    assert lib.v2v_comm_check(comm, None) == 0
    assert np.abs(g.cpu().numpy() - gref).max() <= 1e-5 * np.abs(gref).max()
    assert np.abs(p.cpu().numpy() - rp).max() <= 2e-6
    assert np.array_equal(exo.cpu().numpy(), extra)
    lib.v2v_comm_destroy(comm)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_two_ranks_match_single_process():
    env = dict(os.environ)
    env.pop("V2V_DP_BACKEND", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DP_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
