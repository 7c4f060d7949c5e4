import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import v2v_gnn_b200 as v2v
from oracle import v2v_oracle as O, torch_ref as T
G='/root/repo/tests/golden'
for name in sorted(f[:-4] for f in os.listdir(G) if f.endswith('.npz')):
    z=np.load(os.path.join(G,name+'.npz'))
    N,S,ps=int(z['N']),int(z['S']),bool(z['per_slot'])
    d=O.BrainDims(N,stages=S,per_slot=ps)
    brain=v2v.BS(N,3,1,16,1,4,stages=S,per_slot=ps,max_batch=64,data_parallel=False)
    brain.set_flat_params(z['params'],0)
    x={"Node_Input":z['node'],"Edge_Input":z['edge'],"Adjacency_Matrix":z['adj']}
    q=np.stack(brain.predict(x),1)
    h=brain.train_dnn(x,{"Decide_Output":z['y']},z['node'].shape[0])
    g=brain.get_flat_params(2)
    L=O.unflatten_params(d,z['params'].astype(np.float64))
    # fp32 torch cpu grads
    tl=T.to_torch_layers(L,dtype=torch.float32,requires_grad=True)
    qt=T.forward_factored(d,tl,torch.tensor(z['node']),torch.tensor(z['edge']),torch.tensor(z['adj']))
    lt,_=T.huber_total(qt,torch.tensor(z['y'])); lt.backward()
    g32=np.concatenate([np.concatenate([l['W'].grad.numpy().ravel(),l['b'].grad.numpy().ravel()]) for l in tl])
    gr=z['grads']
    print(name,'|q|max',np.abs(z['q']).max(),'q err',np.abs(q-z['q']).max()/np.abs(z['q']).max(),
          'loss',h.history['loss'][0],float(z['loss']))
    print('   gpu grad relerr',np.abs(g-gr).max()/np.abs(gr).max(),' torch-fp32 grad relerr',np.abs(g32-gr).max()/np.abs(gr).max(),
          ' l2 gpu',np.linalg.norm(g-gr)/np.linalg.norm(gr),' l2 torch32',np.linalg.norm(g32-gr)/np.linalg.norm(gr))
    o=0
    for li,(K,Nout) in enumerate(d.layer_shapes()):
        n=d.G*K*Nout+d.G*Nout
        e=np.abs(g[o:o+n]-gr[o:o+n]).max(); m=np.abs(gr[o:o+n]).max()
        print('     layer',li,'max|g|',m,'abs err',e,'rel',e/max(m,1e-30)); o+=n
