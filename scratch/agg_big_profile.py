import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from bench import synth_numpy
lib = v2v.load_library(); ptr = v2v._lib.ptr
for N, B in ((32, 8192), (64, 8192), (128, 8192)):
    rng = np.random.default_rng(N)
    _, _, adj = synth_numpy(256, N, rng)
    im0, _, _ = v2v.pack_adjacency(torch.from_numpy(adj).cuda())
    im = im0.repeat((B // 256, 1, 1)).contiguous()
    H = torch.randn(B, N, 16, device="cuda"); out = torch.empty_like(H)
    for _ in range(3):
        v2v._lib.check(lib.v2v_agg_mask_ex(ptr(H), ptr(im), None, ptr(out), B, N, 16, 0, 1, v2v._lib.current_stream()))
    torch.cuda.synchronize()
