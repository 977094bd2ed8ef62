#!/bin/bash
mkdir -p gpurun_out
export GSG_RT_BUDGET_KB=60
timeout 110 compute-sanitizer --tool racecheck python tools/sanitize_r2.py > gpurun_out/san_racecheck2.txt 2>&1; tail -n 4 gpurun_out/san_racecheck2.txt
unset GSG_RT_BUDGET_KB
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/san2_bench.json 2> gpurun_out/san2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/san2_bench.json').read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
PY
