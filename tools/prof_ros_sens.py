"""One launch of k_rosenbrock23_sens on the robertson training shape (np = 43, 65 536 trajectories) for ncu."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from crnn_b200 import cases, synth
from crnn_b200.engine import Engine
from problems import make_problem
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "checkpoints.json")))
eng = Engine(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pb = make_problem("robertson", golden, 512)
c = pb["case"]
u0 = synth.make_u0("robertson", N)
truth = eng.solve_batch(cases.true_model_robertson(), c.opts(pred_clamp=(-np.inf, np.inf)), u0, want_stats=False)["pred"]
ud = torch.from_numpy(u0).cuda(); dd = torch.from_numpy(truth * 1.0001).cuda()
for _ in range(2):
    r = eng.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], ud, dd, pb["yscale"], c.loss_kind, want_stats=False)
torch.cuda.synchronize()
print("ok", int((r["retcode"] == 1).sum()))
