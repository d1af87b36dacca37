import sys, json
sys.path.insert(0,'.')
import numpy as np
from oracle import oracle
from crnn_b200 import cases, _abi
from crnn_b200.engine import Engine
eng=Engine(0)
m=cases.synthetic_stiff_model(); u0=cases.synthetic_stiff_u0(256); o=cases.synthetic_stiff_opts()
got=eng.solve_batch(m,o,u0); ref=oracle.solve_batch(m,o,u0,n_threads=8)
scale=np.maximum(np.abs(ref['pred']).max(axis=(0,1)),1e-4)
err=np.abs(got['pred']-ref['pred'])/scale
print('same counts frac', (got['stats']['n_rhs']==ref['stats']['n_rhs']).mean())
print('final-time err max', err[:,-1].max(), ' all-saves err max', err.max())
i,k,s=np.unravel_index(err.argmax(),err.shape); print('worst traj',i,'save',k,'species',s,'got',got['pred'][i,k,s],'ref',ref['pred'][i,k,s], 'stats got',got['stats'][i],'ref',ref['stats'][i])
print('err per save (max over traj,species):', np.round(err.max(axis=(0,2)),4))
# tight tolerance: do both converge to the same thing?
ot=cases.synthetic_stiff_opts(); ot.abstol=1e-12; ot.reltol=1e-8; ot.maxiters=10**6
g2=eng.solve_batch(m,ot,u0[:32]); r2=oracle.solve_batch(m,ot,u0[:32],n_threads=8)
e2=np.abs(g2['pred']-r2['pred'])/scale
print('tight: final err', e2[:,-1].max(), 'all', e2.max(), 'steps', g2['stats']['n_accept'][:4], r2['stats']['n_accept'][:4])
# small n=30 but same dims with Rosenbrock? compare oracle kc4 tight vs loose at final time
print('oracle loose vs tight final', (np.abs(ref['pred'][:32,-1]-r2['pred'][:,-1])/scale).max(), 'gpu loose vs tight final', (np.abs(got['pred'][:32,-1]-g2['pred'][:,-1])/scale).max())
