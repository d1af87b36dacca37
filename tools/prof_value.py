"""One launch of the predict kernels on BASELINE config 3 (robertson, Rosenbrock23, 262 144 ICs) / config 2's shape for ncu.
usage: python tools/prof_value.py robertson|case2 [alg]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from crnn_b200 import cases, synth, _abi
from crnn_b200.engine import Engine
from problems import make_problem
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
name = sys.argv[1] if len(sys.argv) > 1 else "robertson"
eng = Engine(0)
pb = make_problem(name, golden, 64)
N = 262144 if name == "robertson" else 65536
o = pb["opts"] if len(sys.argv) < 3 else pb["case"].opts(obs_idx=pb["opts"].obs_idx, alg=int(sys.argv[2]))
ud = torch.from_numpy(synth.make_u0(name, N)).cuda()
pred = torch.empty((N, o.n_save, pb["case"].ns), dtype=torch.float64, device="cuda")
for _ in range(2):
    r = eng.solve_batch(pb["model"], o, ud, out=pred, want_stats=False)
torch.cuda.synchronize()
print("ok", int((r["retcode"] == 1).sum()))
