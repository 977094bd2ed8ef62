"""Per-CTA time stamps of the row-tile kernel (development aid): where a CTA's lifetime goes and how the SMs fill.
usage: python tools/stamps_rowtile.py d [beta]"""
import ctypes as C, math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsg_b200 as g
from gsg_b200 import lib
D, k, n, d = 6, 3, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 1
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
os.environ["GSG_ONLY_CLASS"] = "1"
plan = g.Plan(D, k, n)
print(plan.describe().splitlines()[0])
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
buf = np.zeros(8192, dtype=np.int64)
lib.gsg_debug_stamps(plan._h, buf.ctypes.data_as(C.c_void_p), 8192)
b = buf.reshape(1024, 8)
b = b[b[:, 0] > 0]
t0 = b[:, 0].min()
start, ready, end = (b[:, 0] - t0) / 1e3, (b[:, 1] - t0) / 1e3, (b[:, 2] - t0) / 1e3
main_max, smid, main_sum, epi_sum = b[:, 3] / 1965.0, b[:, 4], b[:, 5] / 1965.0 / 8, b[:, 6] / 1965.0 / 8
print(f"{len(b)} CTAs on {len(set(smid.tolist()))} SMs, kernel span {end.max():.1f} us")
print(f"per CTA (us): load wait {np.mean(ready - start):.2f}  main (slowest warp) {np.mean(main_max):.2f}  main (mean warp) {np.mean(main_sum):.2f}  "
      f"epilogues (mean warp) {np.mean(epi_sum):.2f}  total {np.mean(end - start):.2f}")
gaps, busy = [], []
for s in sorted(set(smid.tolist())):
    idx = np.where(smid == s)[0]
    idx = idx[np.argsort(start[idx])]
    busy.append(sum(end[i] - start[i] for i in idx))
    gaps += [start[idx[j + 1]] - end[idx[j]] for j in range(len(idx) - 1)]
print(f"per SM: CTAs {len(b) / len(busy):.2f}, busy {np.mean(busy):.1f} us (max {np.max(busy):.1f}), gap between CTAs {np.mean(gaps):.2f} us, first start {np.mean([start[smid == s].min() for s in set(smid.tolist())]):.2f}, last end mean {np.mean([end[smid == s].max() for s in set(smid.tolist())]):.1f}")
tile = b[:, 7]
for t in sorted(set(tile.tolist()))[:40]:
    m = tile == t
    print(f"  tile {t:3d}: {m.sum():4d} CTAs  load {np.mean((ready - start)[m]):5.2f}  main max {np.mean(main_max[m]):5.2f} mean {np.mean(main_sum[m]):5.2f}  epi {np.mean(epi_sum[m]):5.2f}  total {np.mean((end - start)[m]):5.2f}")
