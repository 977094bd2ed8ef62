#!/bin/bash
# call 12: persistent row-tile kernel (run-ahead fix), stamps, bench, DFMA ILP microbenchmark
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
timeout 60 python tools/stamps_rowtile.py 1 > gpurun_out/c13_stamps_d1.txt 2>&1; head -12 gpurun_out/c13_stamps_d1.txt
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c13_bench.json').read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
PY
