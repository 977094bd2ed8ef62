"""Run a few fused gradient applies (for ncu): python tools/one_grad.py D k n reps"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
D, k, n, reps = [int(a) for a in sys.argv[1:5]]
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
a = -np.ones(D)
for _ in range(reps):
    plan.apply_grad_dev(a, x, y)
plan.sync()
torch.cuda.synchronize()
