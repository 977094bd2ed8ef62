#!/bin/bash
# call 6: smem bank microbenchmark, full GPU suite with the row-tile kernel as the default, bench, per-direction ncu counters
mkdir -p gpurun_out
./tools/lds_bench > gpurun_out/c6_lds.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/c6_pytest.txt; cat gpurun_out/c6_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
tail -c 400 gpurun_out/c6_bench.json
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
for d in 1 2 3 4 5 6; do
  timeout 200 ncu --metrics $M --clock-control none -k regex:sweep_rowtile -s 1 -c 1 --csv --log-file gpurun_out/c6_rt_d$d.csv python tools/one_apply.py 6 3 8 $d 2 > /dev/null 2>&1
done
for d in 1 4; do
  GSG_RT_C=2 GSG_RT_BUDGET_KB=112 timeout 200 ncu --metrics $M --clock-control none -k regex:sweep_rowtile -s 1 -c 1 --csv --log-file gpurun_out/c6_rt_c2_d$d.csv python tools/one_apply.py 6 3 8 $d 2 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/c6_rt_*d?.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if len(rows)<2: print(f,'EMPTY'); continue
    h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value')
    print(f, {r[mi].split('.')[0][-28:]:r[vi] for r in rows[1:]})
PY
cat gpurun_out/c6_lds.txt
