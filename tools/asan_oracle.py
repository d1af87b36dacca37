"""The CPU oracle under AddressSanitizer + UBSan: every algorithm, RHS flavour and sensitivity mode on small batches.
  gcc -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fopenmp -fPIC -std=gnu11 -ffp-contract=off -shared -o /tmp/liboracle_asan.so oracle/crnn_oracle.c oracle/lean_math_host.c -lm
  LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_oracle.py
(rounds 1 and 2, incl. TRBDF2, F4 and the F4 adjoint: clean)"""
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
dy=oracle.solve_batch(my, cases.yeast_opts(alg=0, n_save=40), uy*1.02)['pred']      # F4 gradients: adj_rhs_f4 in the extended weight space
for sm in (2,3):
    oracle.loss_grad_batch(my, cases.yeast_opts(alg=0, n_save=40, sens_mode=sm), cases.yeast_seed(np.array(golden['yeast']['p'])), uy, dy, np.ones(7), n_threads=2)
from test_f4_mlp_cpu import qssa_like_model
from crnn_b200.model import SolveOpts
q=qssa_like_model(); uq=0.2+np.random.default_rng(3).random((3,3))
for sm in (2,3):
    oq=SolveOpts(saveat=np.linspace(0.0,2.0,21), t0=0.0, t1=2.0, alg=0, abstol=1e-8, reltol=1e-6, maxiters=100000, sens_mode=sm, obs_idx=np.array([0,2]))
    oracle.loss_grad_batch(q, oq, np.eye(q.n_w), uq, np.ones((3,21,2)), np.ones(2), n_threads=2)
print('asan/ubsan run ok')
