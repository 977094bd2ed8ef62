"""Per-warp clock stamps of the register-tiled long kernel's first 8 CTAs (development aid).
usage: GSG_ONLY_CLASS=<i> python tools/stamps_long.py d beta"""
import ctypes as C, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
from gsg_b200 import lib
D, k, n, d = 6, 3, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 1
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
lib.gsg_debug_stamps.restype = C.c_int
lib.gsg_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for _ in range(2):
    plan.apply_D_dev(d, x, y, 1.0, beta)
plan.sync(); torch.cuda.synchronize()
lib.gsg_debug_stamps(plan._h, None, 0)          # enable
plan.apply_D_dev(d, x, y, 1.0, beta)
plan.sync(); torch.cuda.synchronize()
buf = np.zeros(64 * 8, dtype=np.int64)
lib.gsg_debug_stamps(plan._h, buf.ctypes.data_as(C.c_void_p), 64 * 8)
b = buf.reshape(8, 8, 8)
t0 = b[..., 0][b[..., 0] > 0].min() if (b[..., 0] > 0).any() else 0
print("cta warp | issue  stage_wait  records  | nrec nrows | start")
for cta in range(8):
    for w in range(8):
        s, i, st, e, nrec, nrows = b[cta, w, :6]
        if s == 0:
            continue
        print(f"{cta:3d} {w:3d} | {i-s:6d} {st-i:8d} {e-st:8d} | {nrec:4d} {nrows:4d} | {s-t0:8d}  per-rec {(e-st)/max(nrec,1):.0f}")
