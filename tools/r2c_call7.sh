#!/bin/bash
# call 7: row-tile kernel v2 (row groups, bank-permuted lanes, staged bulk output): parity, bench, per-direction counters, one full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 4
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -k "config4 or mg_virtual" 2>&1 | tail -n 4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "launches", d["gpu_launches"], "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
PY
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
for d in 1 3 4 6; do
  timeout 200 ncu --metrics $M --clock-control none -k regex:sweep_rowtile -s 1 -c 1 --csv --log-file gpurun_out/c7_rt_d$d.csv python tools/one_apply.py 6 3 8 $d 2 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/c7_rt_d?.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if len(rows)<2: print(f,'EMPTY'); continue
    h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value')
    print(f, {r[mi].split('.')[0][-28:]:r[vi] for r in rows[1:]})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_rowtile -s 1 -c 1 -f -o gpurun_out/rt2_full python tools/one_apply.py 6 3 8 2 2 > gpurun_out/c7_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/rt2_full.ncu-rep > gpurun_out/rt2_full_summary.txt 2>&1; cat gpurun_out/rt2_full_summary.txt
