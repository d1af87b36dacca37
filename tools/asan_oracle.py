"""The CPU oracle under AddressSanitizer + UBSan: every algorithm, RHS flavour and sensitivity mode on small batches.
  gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -fPIC -std=gnu11 -shared -o /tmp/liboracle_asan.so oracle/crnn_oracle.c -lm
  LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_oracle.py
(round 1: clean)"""
import sys, json
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from oracle import oracle
oracle.LIB='/tmp/liboracle_asan.so'; oracle.build=lambda force=False: oracle.LIB
from crnn_b200 import cases, _abi
from problems import make_problem
golden=json.load(open(os.path.join(ROOT, 'tests', 'golden', 'checkpoints.json')))
YS=np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
for name in ('case1','case2','case3','robertson'):
    pb=make_problem(name,golden,6)
    for alg in (0,1,2,3,4,5):
        oracle.solve_batch(pb['model'], pb['case'].opts(alg=alg, obs_idx=pb['opts'].obs_idx), pb['u0'], n_threads=2)
    oracle.loss_grad_batch(pb['model'], pb['opts'], pb['seed'], pb['u0'], pb['data'], pb['yscale'], pb['loss_kind'], n_threads=2, want_pred=True, want_grad_each=True)
    if name!='robertson':
        for sm in (2,3):
            oracle.loss_grad_batch(pb['model'], pb['case'].opts(obs_idx=pb['opts'].obs_idx, sens_mode=sm), pb['seed'], pb['u0'], pb['data'], pb['yscale'], pb['loss_kind'], n_threads=2)
m,seed=cases.hychem_model(cases.hychem_p(0, lnA_shift=-2.0), YS); u0=cases.hychem_u0(4)
for alg in (0,1,2,3,4,5):
    r=oracle.solve_batch(m, cases.hychem_opts(alg=alg), u0, n_threads=2)
for sm in (1,2,3):
    oracle.loss_grad_batch(m, cases.hychem_opts(alg=0, sens_mode=sm), seed, u0, r['pred']*1.01, YS, n_threads=2)
nsu=np.array([3,40,1,17],dtype=np.int32)
oracle.solve_batch(m, cases.hychem_opts(alg=3, maxiters=7), u0, n_save_used=nsu)
my=cases.yeast_model(np.array(golden['yeast']['p']))      # F4: MLP-augmented inputs, finite-difference Jacobian
uy=cases.YEAST_IC_LB+np.random.default_rng(0).random((3,7))*(cases.YEAST_IC_UB-cases.YEAST_IC_LB)
for alg in (0,1,3,4,5):
    oracle.solve_batch(my, cases.yeast_opts(alg=alg, n_save=40), uy, n_threads=2)
print('asan/ubsan run ok')
