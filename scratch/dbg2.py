import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import v2v_gnn_b200 as v2v
from oracle import v2v_oracle as O
for (N,S,B) in [(20,3,6),(20,2,16),(4,3,8)]:
    rng=np.random.default_rng(0)
    brain=v2v.BS(N,3,1,16,1,4,stages=S,per_slot=False,max_batch=64,data_parallel=False,seed=1)
    print(N,S,B,brain.fused_info(B,True)); sys.stdout.flush()
    node,edge,adj,_=O.synth_batch(B,N,rng)
    nd,ed,ad=(torch.from_numpy(t.astype(np.float32)).cuda() for t in (node,edge,adj))
    im,om,_=v2v.pack_adjacency(ad)
    q=brain.forward_device(nd,ed,in_mask=im); torch.cuda.synchronize(); print(' fwd ok'); sys.stdout.flush()
    y=q+1
    hl=brain.train_step_device(nd,ed,im,om,None,y); torch.cuda.synchronize(); print(' train ok',hl.sum().item()); sys.stdout.flush()
