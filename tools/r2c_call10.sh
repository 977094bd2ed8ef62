#!/bin/bash
# call 10: software-pipelined main loop + persistent CTAs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
GSG_RT_GRID=0 timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
python tools/stamps_rowtile.py 1 > gpurun_out/c10_stamps_d1.txt 2>&1; head -12 gpurun_out/c10_stamps_d1.txt
GSG_RT_GRID=0 python tools/stamps_rowtile.py 1 > gpurun_out/c10_stamps_d1_np.txt 2>&1; head -12 gpurun_out/c10_stamps_d1_np.txt
for v in "GSG_RT_GRID=-1" "GSG_RT_GRID=0"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c10_bench.json').read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
PY
done
