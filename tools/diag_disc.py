import sys, json
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from oracle import oracle
from crnn_b200 import _abi
from crnn_b200.engine import Engine
from problems import make_problem
g=json.load(open('tests/golden/checkpoints.json'))
eng=Engine(0)
for name,N in [('case3',128),('case2',64)]:
    pb=make_problem(name,g,N); c=pb['case']
    data=np.abs(pb['data'])+1e-6 if name=='case3' else pb['data']
    for ragged in (False,True):
        nsu=np.random.default_rng(5).integers(1,c.n_save+1,size=N).astype(np.int32) if ragged else None
        od=c.opts(obs_idx=np.arange(c.ns),sens_mode=_abi.SENS_DISCRETE_ADJOINT); of=c.opts(obs_idx=np.arange(c.ns),err_norm_includes_sens=False)
        args=(pb['seed'],pb['u0'],data,pb['yscale'],pb['loss_kind'])
        got=eng.loss_grad_batch(pb['model'],od,*args,n_save_used=nsu); ref=oracle.loss_grad_batch(pb['model'],od,*args,n_save_used=nsu,n_threads=8)
        fwd=eng.loss_grad_batch(pb['model'],of,*args,n_save_used=nsu); rfw=oracle.loss_grad_batch(pb['model'],of,*args,n_save_used=nsu,n_threads=8)
        def rel(a,b): return np.abs(a-b).max()/np.abs(b).max()
        print(name,'ragged',ragged,'| disc gpu-vs-oracle loss',rel(got['loss'],ref['loss']),'| fwd gpu-vs-oracle loss',rel(fwd['loss'],rfw['loss']),'| disc-vs-fwd oracle',rel(ref['loss'],rfw['loss']))
        bad=np.nonzero(np.abs(fwd['loss']-rfw['loss'])>1e-6*np.abs(rfw['loss']))[0]
        if bad.size: print('   fwd bad idx',bad[:10],'nsu',None if nsu is None else nsu[bad[:10]], fwd['loss'][bad[:4]], rfw['loss'][bad[:4]], 'n_saved', fwd['n_saved'][bad[:4]], rfw['n_saved'][bad[:4]])
        bad=np.nonzero(np.abs(got['loss']-ref['loss'])>1e-6*np.abs(ref['loss']))[0]
        if bad.size: print('   disc bad idx',bad[:10],'nsu',None if nsu is None else nsu[bad[:10]], got['loss'][bad[:4]], ref['loss'][bad[:4]])
