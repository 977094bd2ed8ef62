#!/bin/bash
mkdir -p gpurun_out
export GSG_ROWTILE=1
ncu --set full --clock-control none --import-source on -k regex:sweep_rowtile -s 8 -c 1 -f -o gpurun_out/rt_full \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rt.log 2>&1
python tools/ncu_summary.py gpurun_out/rt_full.ncu-rep > gpurun_out/rt_full_summary.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/rt_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/rt_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:61]:
    print(r[ki][:60], r[vi])
PY
