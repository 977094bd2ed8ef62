"""Timeline of one sweep: per kernel class, when and on how many SMs its CTAs ran (development aid)."""
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
if os.environ.get("GSG_PART"):
    r_, w_ = [int(t) for t in os.environ["GSG_PART"].split(",")]
    plan.set_partition(r_, w_)
lib.gsg_debug_stamps.restype = C.c_int
lib.gsg_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for _ in range(3):
    plan.apply_D_dev(d, x, y, 1.0, beta)
plan.sync(); torch.cuda.synchronize()
lib.gsg_debug_stamps(plan._h, None, 0)
plan.apply_D_dev(d, x, y, 1.0, beta)
plan.sync(); torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.int64)
lib.gsg_debug_stamps(plan._h, buf.ctypes.data_as(C.c_void_p), 8192)
ch = buf[1024:1024 + 2048].reshape(512, 4); ch = ch[ch[:, 1] > 0]
st = buf[4096:4096 + 2048].reshape(512, 4); st = st[st[:, 1] > 0]
lg = buf[6144:6144 + 2000].reshape(500, 4); lg = lg[lg[:, 1] > 0]
t0 = min([a[:, 1].min() for a in (ch, st, lg) if len(a)])
def show(name, a):
    if not len(a): return
    s = (a[:, 1] - t0) / 1e3; e = (a[:, 2] - t0) / 1e3
    print(f"{name:14s} CTAs {len(a):4d} SMs {len(set(a[:,0])):3d} start {s.min():6.1f}..{s.max():6.1f} end {e.min():6.1f}..{e.max():6.1f}  mean life {np.mean(e-s):6.1f} us  SM-time {np.sum(e-s)/148:6.1f} us-equiv")
show("const-H p=4", ch)
show("stream", st)
for tag in sorted(set(lg[:, 3])):
    show(f"long2 p={tag//100} pass {tag%100}", lg[lg[:, 3] == tag])
print("sweep end:", max([((a[:, 2] - t0) / 1e3).max() for a in (ch, st, lg) if len(a)]))
