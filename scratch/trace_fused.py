import sys, numpy as np, torch, ctypes as C
sys.path.insert(0,'/root/repo')
import v2v_gnn_b200 as v2v
from bench import synth_numpy
N=int(sys.argv[3]) if len(sys.argv)>3 else 20; PS=len(sys.argv)>4 and sys.argv[4]=="slot"; S=int(sys.argv[1]) if len(sys.argv)>1 else 2; B=int(sys.argv[2]) if len(sys.argv)>2 else 1024
rng=np.random.default_rng(0)
brain=v2v.BS(N,3,1,16,1,4,stages=S,per_slot=PS,max_batch=B,data_parallel=False,seed=1)
print(brain.fused_info(B,True))
node,edge,adj=synth_numpy(B,N,rng)
nd,ed,ad=(torch.from_numpy(t.astype(np.float32)).cuda() for t in (node,edge,adj))
im,om,_=v2v.pack_adjacency(ad)
q=brain.forward_device(nd,ed,in_mask=im); y=q+1
for _ in range(3): brain.train_step_device(nd,ed,im,om,None,y)
for _ in range(3): brain.train_step_device(nd,ed,im,om,None,y)
buf=torch.zeros(49*12*2,dtype=torch.int64,device='cuda')
lib=v2v.load_library()
assert lib.v2v_fused_set_trace(C.c_void_p(buf.data_ptr()))==0
brain.train_step_device(nd,ed,im,om,None,y); torch.cuda.synchronize()
lib.v2v_fused_set_trace(None)
t=buf.cpu().numpy().reshape(49,12,2)
nph=brain.fused_info(B,True)['phases']
names=[]
for s in range(S): names+= [f'gemm stage{s}',f'agg{s}']
names+=['mlp1 41x80','mlp2 80x40','mlp3 40x20','mlp4 20x4','loss','bwd mlp4','bwd mlp3','bwd mlp2','bwd mlp1','aggT']
for s in range(S-1,0,-1): names+=[f'bwd stage{s}','aggT']
names+=['bwd stage0']
t0=t[0,:,1].min()
tot=t[nph,:,1].max()-t0
print('phase            wall   | per-warp busy cycles (work end - phase start)')
for i in range(nph):
    start=t[i,:,1].max(); end=t[i+1,:,1].max()
    busy=(t[i+1,:,0]-t[i,:,1])
    print(f'{names[i] if i<len(names) else i:14s} {end-start:7d} {100*(end-start)/tot:5.1f}% | '+' '.join(f'{b:6d}' for b in busy))
print('total',tot,'cycles')
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): brain.train_step_device(nd,ed,im,om,None,y)
e1.record(); torch.cuda.synchronize()
print('avg step us',1e3*e0.elapsed_time(e1)/50)
e0.record()
for _ in range(50): brain.forward_device(nd,ed,in_mask=im,out=q)
e1.record(); torch.cuda.synchronize()
print('avg fwd us',1e3*e0.elapsed_time(e1)/50)
