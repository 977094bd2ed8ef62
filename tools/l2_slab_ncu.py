"""Three back-to-back sweeps (d = 1, 2, 3) restricted to a slab via the block partition; run under
ncu --cache-control none to see how much of the slab the 2nd / 3rd sweep still find in L2.
usage: l2_slab_ncu.py rank world"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
rank, world = int(sys.argv[1]), int(sys.argv[2])
D, k, n = 6, 3, 8
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
plan.set_partition(rank, world)
offs, sizes, _ = plan.partition_blocks(0)
print("slab MB", sizes.sum() * 8 / 1e6, flush=True)
flush = torch.empty(80_000_000, dtype=torch.float64, device="cuda")
for rep in range(2):
    flush.fill_(1.0); torch.cuda.synchronize()
    plan.apply_D_dev(1, x, y, 1.0, 0.0)
    plan.apply_D_dev(2, x, y, 1.0, 1.0)
    plan.apply_D_dev(3, x, y, 1.0, 1.0)
    plan.sync(); torch.cuda.synchronize()
