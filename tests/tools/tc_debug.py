import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import v2v_gnn_b200 as v2v
from oracle import v2v_oracle as O
np.set_printoptions(precision=4, suppress=True, linewidth=200)
N, S, B = 20, 2, 6
rng = np.random.default_rng(1)
d = O.BrainDims(N, stages=S, per_slot=False)
L = O.init_params(d, rng, bias_scale=0.05)
node, edge, adj, _ = O.synth_batch(B, N, rng)
node, edge = node.astype(np.float32), edge.astype(np.float32)
brain = v2v.BS(N, 3, 1, 16, 1, 4, stages=S, per_slot=False, max_batch=B, data_parallel=False)
brain.set_flat_params(O.flatten_params(L), 0)
lib = brain._lib
dev = lambda a: torch.from_numpy(a).cuda()
nd, ed = dev(node), dev(edge)
im, _, _ = v2v.pack_adjacency(dev(adj.astype(np.float32)))
q = torch.zeros(B, N, 4, device="cuda")
# layer 0 reference: [node | edge] @ [W1; W2] (neighbour rows skipped)
W0 = L[0]["W"][0] if L[0]["W"].ndim == 3 else L[0]["W"]
print("W0 shape", W0.shape)
x = np.concatenate([node.reshape(-1, 9), edge.reshape(-1, 4)], 1).astype(np.float64)
ref0 = x @ W0[:13]
for layer in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    npad = C.c_int32()
    dbg = torch.full((128, 128), float("nan"), device="cuda")
    v2v._lib.check(lib.v2v_brain_tc_debug(brain._handle, nd.data_ptr(), ed.data_ptr(), im.data_ptr(), B, layer, q.data_ptr(),
                                          dbg.data_ptr(), C.byref(npad), None))
    torch.cuda.synchronize()
    acc = dbg.cpu().numpy().reshape(-1)[:128 * npad.value].reshape(128, npad.value)
    print("layer", layer, "Npad", npad.value, "nan count", np.isnan(acc).sum())
    if layer == 0:
        print("acc[:4,:16]\n", acc[:4, :16]); print("ref[:4]\n", ref0[:4])
        print("max abs err rows<120:", np.abs(acc[:120, :16] - ref0[:120]).max(), " ref scale", np.abs(ref0).max())
        # diagnostics: does acc match ref with some permutation?
        print("col sums acc", acc[:120].sum(0)); print("col sums ref", ref0[:120].sum(0))
        print("row sums acc", acc[:8].sum(1)); print("row sums ref", ref0[:8].sum(1))
    else:
        print(acc[:3, :min(16, npad.value)])
