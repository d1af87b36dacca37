import sys, os, json, time, subprocess
sys.path.insert(0,'.')
import numpy as np, torch
from crnn_b200 import cases, synth
from crnn_b200.engine import Engine
import bench
eng=Engine(0)
c, model, seed, opts, u0_h, data_h, yscale = bench.build_inputs(eng, 0)
u0_p=torch.from_numpy(u0_h).pin_memory().numpy(); data_p=torch.from_numpy(data_h).pin_memory().numpy()
# raw H2D bandwidth
d=torch.empty(data_h.shape,dtype=torch.float64,device='cuda'); tp=torch.from_numpy(data_p)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): d.copy_(tp,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5; print('H2D GB/s', data_h.nbytes/dt/1e9, 'ms', dt*1e3)
for ramp, ch in [(r, c_) for r in (0, 4, 8) for c_ in (2,3,4,5,6,8)]:
    os.environ['CRNN_B200_CHUNKS']=str(ch); os.environ['CRNN_B200_RAMP']=str(ramp)
    for _ in range(2): eng.loss_grad_batch(model,opts,seed,u0_p,data_p,yscale,c.loss_kind,want_stats=False)
    t=time.perf_counter()
    for _ in range(10): eng.loss_grad_batch(model,opts,seed,u0_p,data_p,yscale,c.loss_kind,want_stats=False)
    dt=(time.perf_counter()-t)/10
    print('ramp',ramp,'chunks',ch,'ms',round(dt*1e3,2),'traj/s',round(65536/dt))
