"""reconstruct_DG throughput (BASELINE config 5: D=4 sparse k=4 n=8) -- points per second, state resident.
usage: python tools/bench_reconstruct.py [npts] [sorted]"""
import math, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
npts = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
D, k, n = 4, 4, 8
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
u0 = g.tensor_construct(D, k, n, [v1] * D)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); plan.set_stream(stream)
coef = plan.to_device(u0)
gen = torch.Generator(device=dev); gen.manual_seed(20240)
pts = torch.rand(npts, D, dtype=torch.float64, device=dev, generator=gen)
out = torch.empty(npts, dtype=torch.float64, device=dev)
res = {}
for label in ("random", "sorted"):
    if label == "sorted":      # Morton-like key: interleave the top 7 bits of every coordinate
        q = (pts * 128).to(torch.int64).clamp_(0, 127)
        key = torch.zeros(npts, dtype=torch.int64, device=dev)
        for b in range(6, -1, -1):
            for d in range(D):
                key = (key << 1) | ((q[:, d] >> b) & 1)
        pts = pts[torch.argsort(key)].contiguous()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.reconstruct_dev(coef, pts, npts, out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    exact = torch.prod(torch.sin(2 * math.pi * pts), dim=1)
    err = float((out - exact).abs().max())
    res[label] = {"ms": best, "points_per_s": npts / best * 1e3, "max_abs_err_vs_exact": err}
    print(label, res[label], flush=True)
print(json.dumps({"npts": npts, "D": D, "k": k, "n": n, "N": plan.size, **res}))
