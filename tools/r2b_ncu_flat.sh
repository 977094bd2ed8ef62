#!/bin/bash
mkdir -p gpurun_out
for cfg in d4k3n7 d2k3n8; do
ncu --set full --clock-control none --import-source on -k regex:sweep_flat -s 12 -c 1 -f -o gpurun_out/flat_$cfg \
   python bench.py --config $cfg --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_flat_$cfg.log 2>&1
python tools/ncu_summary.py gpurun_out/flat_$cfg.ncu-rep > gpurun_out/flat_${cfg}_summary.txt 2>&1
done
tail -n 3 gpurun_out/ncu_flat_d4k3n7.log
