"""A/B of RK4 driver variants at D=6,k=3,n=8 (development aid): python tools/rk4_ab.py"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
D, k, n = 6, 3, 8
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
u0 = g.tensor_construct(D, k, n, [v1] * D)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); plan.set_stream(stream)
x = plan.to_device(u0)
a = np.ones(D)
for mode in (0, 1):
    plan.set_rk4_mode(mode)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.rk4_advect_dev(a, x, 1e-4, 8); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 8)
    print(f"mode {mode}: {best:.3f} ms/step", flush=True)
