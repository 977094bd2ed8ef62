"""Regenerate profiles/traffic.json (bench.py's `roofline.traffic`) from an `ncu --set full` capture of the streaming
kernel family of ONE right-hand side:  python tools/make_traffic_json.py gpurun_out/r2_stream_full.ncu-rep
(capture command: tools/final_n1.sh -- ncu --set full --clock-control none -k regex:sweep_stream -s 9 -c 9 on
tools/one_grad.py 6 3 8 2)."""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
idx = {h: i for i, h in enumerate(rows[0])}
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in rows[2:]:
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    launches.append({"kernel": r[idx["Kernel Name"]].split("(")[0], "block": r[idx["Block Size"]],
                     "registers": int(r[idx["launch__registers_per_thread"]]), "us": float(r[idx["gpu__time_duration.sum"]]),
                     "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6,
                     "dram_GBs": (rd + wr) / (float(r[idx["gpu__time_duration.sum"]]) * 1e-6) / 1e9,
                     "fp64_pipe_pct": float(r[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]])})
avg = sum(1e6 * (l["dram_read_MB"] + l["dram_write_MB"]) for l in launches) / len(launches)
doc = {"sweep_stream_kernel_bytes_per_launch": avg, "launches": launches,
       "source": f"{os.path.basename(rep)}: ncu --set full --clock-control none -k regex:sweep_stream -s 9 -c 9 on tools/one_grad.py 6 3 8 2 "
                 "= the 9 launches of the sweep_stream_kernel family of ONE right-hand side (3 PAIR launches + 6 reduced sweeps; "
                 "the first PAIR launch and the first reduced sweep write with beta = 0, the rest reduce-add); "
                 "dram__bytes_read.sum + dram__bytes_write.sum averaged over the 9 launches; captured on the round-2 HEAD",
       "algorithmic_bytes_per_rhs": 16.0 * 6 * 34455456 * 0.83}
json.dump(doc, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps({k: doc[k] for k in ("sweep_stream_kernel_bytes_per_launch",)}))
for l in launches:
    print(l)
