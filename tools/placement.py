"""Which SMs / when the CTAs of the constant-bank kernel and of the streaming kernel ran (development aid)."""
import ctypes as C, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
from gsg_b200 import lib
D, k, n, d = 6, 3, 8, 1
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
lib.gsg_debug_stamps.restype = C.c_int
lib.gsg_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
for _ in range(2):
    plan.apply_D_dev(d, x, y, 1.0, 0.0)
plan.sync(); torch.cuda.synchronize()
lib.gsg_debug_stamps(plan._h, None, 0)
plan.apply_D_dev(d, x, y, 1.0, 0.0)
plan.sync(); torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.int64)
lib.gsg_debug_stamps(plan._h, buf.ctypes.data_as(C.c_void_p), 8192)
ch = buf[1024:1024 + 2048].reshape(512, 4)
st = buf[4096:4096 + 2048].reshape(512, 4)
ch = ch[ch[:, 1] > 0]; st = st[st[:, 1] > 0]
t0 = min(ch[:, 1].min() if len(ch) else 1 << 62, st[:, 1].min() if len(st) else 1 << 62)
print(f"const-H CTAs {len(ch)}: start {((ch[:,1]-t0)/1e3).min():.1f}..{((ch[:,1]-t0)/1e3).max():.1f} us, end {((ch[:,2]-t0)/1e3).min():.1f}..{((ch[:,2]-t0)/1e3).max():.1f} us, SMs {len(set(ch[:,0]))}")
print(f"stream  CTAs {len(st)}: start {((st[:,1]-t0)/1e3).min():.1f}..{((st[:,1]-t0)/1e3).max():.1f} us, end {((st[:,2]-t0)/1e3).min():.1f}..{((st[:,2]-t0)/1e3).max():.1f} us, SMs {len(set(st[:,0]))}, tiles/CTA {st[:,3].min()}..{st[:,3].max()}")
# overlap: for each stream CTA, was a const-H CTA alive on the same SM at its start?
ov = 0
for s in st:
    m = ch[ch[:, 0] == s[0]]
    if len(m) and ((m[:, 1] < s[1]) & (m[:, 2] > s[1])).any():
        ov += 1
print("stream CTAs that started while a const-H CTA was resident on the same SM:", ov)
hist = np.histogram((st[:, 1] - t0) / 1e3, bins=10)
print("stream start histogram (us):", hist)
