#!/bin/bash
mkdir -p gpurun_out
for cfg in d2k3n8 d4k3n7; do
  timeout 100 python bench.py --config $cfg --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_$cfg.json 2> gpurun_out/r2_bench_$cfg.err
  python - gpurun_out/r2_bench_$cfg.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],5), "staged", d.get("staged_ms_per_step"), "value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"] and (d["roofline"]["kernel"], round(d["roofline"]["avg_launch_ms"],5)))
except Exception as e:
    print(sys.argv[1], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
