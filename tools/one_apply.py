"""Run a few sweeps (for ncu): python tools/one_apply.py D k n d reps"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
D, k, n, d, reps = [int(a) for a in sys.argv[1:6]]
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
for _ in range(reps):
    plan.apply_D_dev(d, x, y, 1.0, 0.0)
plan.sync()
torch.cuda.synchronize()
