#!/bin/bash
# call 14: SM split between the (compute-bound) row-tile launches and the (DRAM-bound) streaming launches
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c14_$name.json 2> gpurun_out/c14_$name.err
  python - gpurun_out/c14_$name.json "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4), d["roofline"]["frac"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
}
run base A=1
run r48s100 GSG_RT_GRID=48 GSG_STREAM_GRID=100
run r64s84 GSG_RT_GRID=64 GSG_STREAM_GRID=84
run r74s74 GSG_RT_GRID=74 GSG_STREAM_GRID=74
run r36s112 GSG_RT_GRID=36 GSG_STREAM_GRID=112
run r148s112 GSG_STREAM_GRID=112
