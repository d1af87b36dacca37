"""One launch of k_wide_solve on the HyChem F2 model (AutoTsit5, 65 536 trajectories) for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crnn_b200 import cases
from crnn_b200.engine import Engine
YS = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
eng = Engine(0)
m, _ = cases.hychem_model(cases.hychem_p(0), YS)
u0 = torch.from_numpy(cases.hychem_u0(65536)).cuda()
for _ in range(2):
    r = eng.solve_batch(m, cases.hychem_opts(), u0, want_stats=False)
torch.cuda.synchronize()
print("ok", int((r["retcode"] == 1).sum()))
