"""Experiment: do sweeps d=1..5 restricted to one l_6-slab stay L2-resident? (development aid)"""
import math, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
D, k, n = 6, 3, 8
L = int(os.environ.get("GSG_SLAB_L6", "-1"))
plan = g.Plan(D, k, n)
N = plan.size
# DOFs of the slab
tot = 0
for lv in g._levels(D, n, "sparse"):
    if L < 0 or lv[-1] - 1 == L:
        tot += int(np.prod([1 << max(0, l - 2) for l in lv])) * k ** D
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); plan.set_stream(stream)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D)); y = torch.zeros_like(x)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
def five():
    for d in range(1, 6):
        plan.apply_D_dev(d, x, y, 1.0, 0.0 if d == 1 else 1.0)
def one():
    plan.apply_D_dev(3, x, y, 1.0, 1.0)
for name, fn, nsw in (("5 sweeps d=1..5", five, 5), ("1 sweep d=3 (cold L2)", one, 1)):
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f"slab L6={L}: {tot} DOFs ({100*tot/N:.1f}% of N, {8e-6*tot:.0f} MB/vector): {name}: {t*1e3:.0f} us "
          f"-> {t*1e6/nsw/(tot/1e6):.2f} ns per MDOF-sweep")
