"""One launch of k_kencarp4_wide on the 30-state synthetic stiff model (BASELINE config 5 shape, 16 384 ICs = one GPU's share) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crnn_b200 import cases
from crnn_b200.engine import Engine
eng = Engine(0)
m = cases.synthetic_stiff_model()
u0 = torch.from_numpy(cases.synthetic_stiff_u0(16384)).cuda()
for _ in range(2):
    r = eng.solve_batch(m, cases.synthetic_stiff_opts(), u0, want_stats=False)
torch.cuda.synchronize()
print("ok", int((r["retcode"] == 1).sum()))
