"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep [regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in idx:
            print(f"{w:90s} {r[idx[w]]} {units[idx[w]]}")
    st = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:7]
    for v, h in st:
        print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f}")
