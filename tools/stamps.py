"""Per-phase clock stamps of the TMA short kernel's CTA 0 (development aid)."""
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
b = buf.reshape(64, 8)
t0 = b[0, 0]
print("it | loader: wait_empty issue | storer: wait_done issue+free | compute: wait_full compute | loader abs")
for it in range(44):
    c0, c1, c2, s0, s1, s2, w, cmp_ = b[it]
    print(f"{it:2d} | {c1-c0:8d} {c2-c1:8d} | {s1-s0:8d} {s2-s1:8d} | {w:8d} {cmp_:8d} | {c0-t0:9d}")
