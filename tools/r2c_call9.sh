#!/bin/bash
# call 9: per-CTA time stamps of the row-tile kernel; 16-warp variant
mkdir -p gpurun_out
python tools/stamps_rowtile.py 1 > gpurun_out/c9_stamps_d1.txt 2>&1; cat gpurun_out/c9_stamps_d1.txt
GSG_RT_C=2 GSG_RT_WARPS=16 python tools/stamps_rowtile.py 1 > gpurun_out/c9_stamps_d1_w16.txt 2>&1; cat gpurun_out/c9_stamps_d1_w16.txt
GSG_RT_C=2 python tools/stamps_rowtile.py 1 > gpurun_out/c9_stamps_d1_c2.txt 2>&1; head -5 gpurun_out/c9_stamps_d1_c2.txt
GSG_RT_WARPS=16 GSG_RT_C=2 timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
for v in "GSG_RT_C=4" "GSG_RT_C=2 GSG_RT_WARPS=16"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c9_bench.json').read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
PY
done
