"""Does a 26 MB slab stay L2-resident across the three local sweeps?  (development aid)
Uses the block partition to restrict the plan to the {level_4 = level_5 = level_6 = 0} slab (rank 7 of 8)."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
D, k, n = 6, 3, 8
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
plan.set_partition(7, 8)
offs, sizes, _ = plan.partition_blocks(0)
print("slab doubles", sizes.sum(), "MB", sizes.sum() * 8 / 1e6)
flush = torch.empty(80_000_000, dtype=torch.float64, device="cuda")     # 640 MB
only = os.environ.get("GSG_CLASS_MASK")
def run(flush_between):
    ts = []
    for rep in range(5):
        flush.fill_(1.0); torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for i, d in enumerate((1, 2, 3)):
            if flush_between:
                flush.fill_(float(rep + i)); torch.cuda.synchronize()
            e[i].record()
            plan.apply_D_dev(d, x, y, 1.0, 0.0 if i == 0 else 1.0)
            if flush_between:
                e[i + 1].record() if i == 2 else None
                ev = torch.cuda.Event(enable_timing=True); ev.record(); torch.cuda.synchronize()
                ts.append((rep, i, e[i].elapsed_time(ev)))
        if not flush_between:
            e[3].record(); torch.cuda.synchronize()
            ts.append((rep, "all3", e[0].elapsed_time(e[3])))
    return ts
a = run(True)
b = run(False)
per = {}
for rep, i, t in a:
    per.setdefault(i, []).append(t)
print("flushed between sweeps: per-sweep ms", {i: round(min(v), 4) for i, v in per.items()}, "sum", round(sum(min(v) for v in per.values()), 4))
print("back to back (slab may stay in L2): three sweeps ms", round(min(t for _, _, t in b), 4))
