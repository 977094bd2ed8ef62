#!/bin/bash
# call 11: persistent CTAs with host-dealt item lists + pair-interleaved main loop
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
python tools/stamps_rowtile.py 1 > gpurun_out/c11_stamps_d1.txt 2>&1; head -12 gpurun_out/c11_stamps_d1.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c11_bench.json').read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_rowtile -s 1 -c 1 -f -o gpurun_out/rt5_full python tools/one_apply.py 6 3 8 1 2 > gpurun_out/c11_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/rt5_full.ncu-rep > gpurun_out/rt5_full_summary.txt 2>&1; head -32 gpurun_out/rt5_full_summary.txt
