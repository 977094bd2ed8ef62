"""Top stalled SASS instructions of one kernel in an .ncu-rep: python tools/ncu_source.py rep kernel-id-filter [topN]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
# the page may hold several kernels: keep the LAST one matching
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k0 = starts[-1]
hdr = rows[k0 + 1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[k0 + 2:] if len(r) == len(hdr) and r[0].startswith("0x")]
rows = rows[k0:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
print("kernel:", rows[0][1][:100], " instructions:", len(body), " samples:", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in body) for h in stalls}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
# cumulative position
cum = 0
ranked = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:top]
for i in sorted(ranked):
    r = body[i]
    s = int(r[idx["# Samples"]] or 0)
    st = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}%  exec={r[idx['Instructions Executed']]:>8s}  {r[idx['Source']].strip()[:70]:70s} {st}")
