import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import torch
from crnn_b200 import _abi, cases
from crnn_b200.engine import Engine, stats_from_torch
from oracle import oracle
eng = Engine(0); dev = torch.device('cuda', 0)
N = 131072
m5 = cases.synthetic_stiff_model(); o5 = cases.synthetic_stiff_opts()
u0 = cases.synthetic_stiff_u0(N); u0_d = torch.from_numpy(u0).to(dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
ms, r = timed(lambda: eng.solve_batch(m5, o5, u0_d, want_stats=False))
print('kc4 131072:', ms, 'ms', N / ms * 1e3, 'traj/s')
M = 16384
got = eng.solve_batch(m5, o5, u0[:M])
with oracle.lu_reciprocal(True), oracle.shared_math(True), oracle.kc4_inverse(True):
    ref = oracle.solve_batch(m5, o5, u0[:M], n_threads=os.cpu_count())
bad = (got['stats']['n_rhs'] != ref['stats']['n_rhs']) | (got['stats']['n_accept'] != ref['stats']['n_accept'])
print('mismatch', bad.sum(), 'of', M, 'retcode ok', (got['retcode'] == 1).all())
scale = np.maximum(np.abs(ref['pred']).max(axis=(0, 1)), 1e-4)
print('err same', (np.abs(got['pred'] - ref['pred'])[~bad] / scale).max(), 'all', (np.abs(got['pred'] - ref['pred']) / scale).max())
