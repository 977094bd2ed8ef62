#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c16_$name.json 2> gpurun_out/c16_$name.err
  python - gpurun_out/c16_$name.json "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"], "share", d["roofline"].get("kernel_share_of_step"))
except Exception as e: print(sys.argv[2], "ERR", e)
PY
}
run rtfirst A=1
run interleave GSG_RT_INTERLEAVE=1
run pool GSG_RT_POOL=1
run rtfirst2 A=1
