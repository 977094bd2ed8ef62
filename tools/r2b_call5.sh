#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
GSG_RT_C=4 timeout 600 python -m pytest tests/test_gpu_rowtile.py -x -q 2>&1 | tail -n 3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c5_$name.json 2> gpurun_out/c5_$name.err
  python - gpurun_out/c5_$name.json "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "launches", d["gpu_launches"], "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4))
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
run c2 GSG_ROWTILE=1
run c2plain GSG_ROWTILE=1 GSG_RT_PLAIN_ORDER=1
run c4 GSG_ROWTILE=1 GSG_RT_C=4
run c4b225 GSG_ROWTILE=1 GSG_RT_C=4 GSG_RT_BUDGET_KB=225
run c1 GSG_ROWTILE=1 GSG_RT_C=1
