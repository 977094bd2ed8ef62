#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" GSG_DESCRIBE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c4_$name.json 2> gpurun_out/c4_$name.err
  python - gpurun_out/c4_$name.json "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "staged", round(d.get("staged_ms_per_step") or 0,4), "launches", d["gpu_launches"], "stream avg ms", d["roofline"] and round(d["roofline"]["avg_launch_ms"],4))
except Exception as e:
    print(sys.argv[2], "ERR", e)
PY
}
run base GSG_ROWTILE=0
run rt112 GSG_ROWTILE=1
run rt225 GSG_ROWTILE=1 GSG_RT_BUDGET_KB=225
run rt75 GSG_ROWTILE=1 GSG_RT_BUDGET_KB=75
head -n 3 gpurun_out/c4_rt112.err
for cfg in d2k3n8 d4k3n7; do
    timeout 300 python bench.py --config $cfg --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/c4_bench_${cfg}.json 2> gpurun_out/c4_bench_${cfg}.err
    python - gpurun_out/c4_bench_${cfg}.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["ms_per_step"], d.get("staged_ms_per_step"), d["roofline"] and d["roofline"]["avg_launch_ms"])
PY
done
