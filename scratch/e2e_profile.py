import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import v2v_gnn_b200 as v2v
from bench import synth_numpy
B,N=1024,20
rng=np.random.default_rng(0)
brain=v2v.BS(N,3,1,16,1,4,stages=2,per_slot=False,max_batch=B,data_parallel=False,seed=1)
node,edge,adj=synth_numpy(B,N,rng)
y=rng.normal(0,1,(B,N,4)).astype(np.float32)
x={"Node_Input":node,"Edge_Input":edge,"Adjacency_Matrix":adj}; yl={"Decide_Output":y}
for _ in range(5): brain.train_dnn(x,yl,B)
def t(f,n=50):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
print('train_dnn total us', t(lambda: brain.train_dnn(x,yl,B)))
print(' _pack_inputs us', t(lambda: brain._pack_inputs(x)))
print(' _pack_labels us', t(lambda: brain._pack_labels(yl,B)))
Bn,nd,ed,ng,ad=brain._pack_inputs(x); yy=brain._pack_labels(yl,B)
print(' _train_rows (C call: H2D+pack+flag sync+step+D2H) us', t(lambda: brain._train_rows(nd,ed,ng,ad,yy,B)))
print('predict total us', t(lambda: brain.predict(x)))
xd={f"D{k+1}_Node_Input":node[:,k].astype(np.float64) for k in range(N)}
xd.update({f"D{k+1}_Edge_Input":edge[:,k].astype(np.float64) for k in range(N)})
xd["Adjacency_Matrix"]=adj.astype(np.float64)
print('predict (reference-style per-slot fp64 dict) us', t(lambda: brain.predict(xd)))
# B=1 acting latency
x1={"Node_Input":node[:1],"Edge_Input":edge[:1],"Adjacency_Matrix":adj[:1]}
print('predict_one_step B=1 us', t(lambda: brain.predict_one_step(x1),200))
b4=v2v.BS(4,3,1,16,1,4,data_parallel=False,seed=1)
n4,e4,a4=synth_numpy(256,4,rng)
x4={f"D{k+1}_Node_Input":n4[:,k].astype(np.float64) for k in range(4)}; x4.update({f"D{k+1}_Edge_Input":e4[:,k].astype(np.float64) for k in range(4)})
x4["Adjacency_Matrix"]=np.kron(a4,np.eye(16))
print('reference-shape N=4 B=256 predict (Kronecker fp64 dict) us', t(lambda: b4.predict(x4),100))
x41={k:v[:1] for k,v in x4.items()}
print('reference-shape N=4 B=1 predict_one_step us', t(lambda: b4.predict_one_step(x41),200))
