"""Quick device-time probe of the sweep kernels at a given (D, k, n): per-direction apply and
RK4 step timings with CUDA events on the plan's stream (development aid, not the bench)."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=6)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
D, k, n = args.D, args.k, args.n

t0 = time.time()
plan = g.Plan(D, k, n)
N = plan.size
print(f"plan D={D} k={k} n={n} N={N} built in {time.time()-t0:.2f}s", flush=True)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
t0 = time.time()
u0 = g.tensor_construct(D, k, n, [v1] * D)
print(f"tensor_construct {time.time()-t0:.2f}s", flush=True)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
plan.set_stream(stream)
x = plan.to_device(u0)
y = torch.zeros_like(x)


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


res = {}
for d in range(1, D + 1):
    best, med = timeit(lambda: plan.apply_D_dev(d, x, y, 1.0, 0.0), args.reps)
    res[f"D{d}_beta0_ms"] = best
    print(f"apply D_{d} beta=0: best {best:.3f} ms  ({16*N/best/1e6:.0f} GB/s algorithmic 16B/DOF)", flush=True)
    best, med = timeit(lambda: plan.apply_D_dev(d, x, y, 1.0, 1.0), args.reps)
    res[f"D{d}_beta1_ms"] = best
    print(f"apply D_{d} beta=1: best {best:.3f} ms  ({24*N/best/1e6:.0f} GB/s actual 24B/DOF)", flush=True)

a = np.ones(D)
best, med = timeit(lambda: plan.rk4_advect_dev(a, x, 1e-4, args.steps), 3)
per = best / args.steps
print(f"RK4: {per:.3f} ms/step -> {N/per*1e3:.3e} DOF-updates/s; model 528B/DOF -> {(64*D+144)*N/per/1e6:.0f} GB/s", flush=True)
res["rk4_ms_per_step"] = per
res["N"] = N
print(json.dumps(res))
