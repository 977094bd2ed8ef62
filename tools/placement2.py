"""Bisect SM sharing: real const-H / real stream kernel against a spinner (development aid).
usage: placement2.py mode   (mode: spinA = real const-H (aux) + spinner 320thr/142KB on main;
                                   spinB = spinner 128thr/50KB on aux + real stream kernel on main)"""
import ctypes as C, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1]
os.environ["GSG_CLASS_MASK"] = "32" if mode in ("spinA", "spinC", "spinD") else "1"
import gsg_b200 as g
from gsg_b200 import lib
D, k, n, d = 6, 3, 8, 1
plan = g.Plan(D, k, n)
v1 = g.vcoeffs_DG(1, k, n, lambda x: math.sin(2 * math.pi * x))
x = plan.to_device(g.tensor_construct(D, k, n, [v1] * D))
y = torch.zeros_like(x)
lib.gsg_debug_stamps.restype = C.c_int
lib.gsg_debug_stamps.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lib.gsg_debug_spin.restype = C.c_int
lib.gsg_debug_spin.argtypes = [C.c_void_p] + [C.c_int] * 6
for _ in range(2):
    plan.apply_D_dev(d, x, y, 1.0, 0.0)
plan.sync(); torch.cuda.synchronize()
lib.gsg_debug_stamps(plan._h, None, 0)
for rep in range(2):
    if mode == "spinA":
        plan.apply_D_dev(d, x, y, 1.0, 0.0)                                   # real const-H on an aux stream
        lib.gsg_debug_spin(plan._h, 0, 148, 320, 141964, 60000, 6144)        # spinner on the main stream
    elif mode == "spinC":
        lib.gsg_debug_spin(plan._h, 1, 148, 320, 141964, 60000, 6144)        # big spinner first, on an aux stream
        plan.apply_D_dev(d, x, y, 1.0, 0.0)                                   # real const-H on another aux stream
    elif mode == "spinD":
        lib.gsg_debug_spin(plan._h, 1, 148, 128, 50176, 60000, 6144)         # small spinner first
        plan.apply_D_dev(d, x, y, 1.0, 0.0)
    else:
        lib.gsg_debug_spin(plan._h, 1, 148, 128, 50176, 40000, 6144)         # spinner on an aux stream
        plan.apply_D_dev(d, x, y, 1.0, 0.0)                                   # real stream kernel on main
    plan.sync(); torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.int64)
lib.gsg_debug_stamps(plan._h, buf.ctypes.data_as(C.c_void_p), 8192)
real = buf[1024:1024 + 2048].reshape(512, 4) if mode in ("spinA", "spinC", "spinD") else buf[4096:4096 + 2048].reshape(512, 4)
spin = buf[6144:6144 + 2048].reshape(512, 4)
real = real[real[:, 1] > 0]; spin = spin[spin[:, 1] > 0]
t0 = min(real[:, 1].min(), spin[:, 1].min())
print(f"{mode}: real CTAs {len(real)} start {((real[:,1]-t0)/1e3).min():.1f}..{((real[:,1]-t0)/1e3).max():.1f} end {((real[:,2]-t0)/1e3).max():.1f} us | "
      f"spinner CTAs {len(spin)} start {((spin[:,1]-t0)/1e3).min():.1f}..{((spin[:,1]-t0)/1e3).max():.1f} end {((spin[:,2]-t0)/1e3).max():.1f} us")
ov = 0
for s in spin:
    m = real[real[:, 0] == s[0]]
    if len(m) and ((m[:, 1] < s[2]) & (m[:, 2] > s[1])).any():
        ov += 1
print("spinner CTAs overlapping in time with a real CTA on the same SM:", ov, "of", len(spin))
