#!/bin/bash
# flat path: CUDA-graph replay of the RK4 step (default) against eager launches
mkdir -p gpurun_out
run() { name=$1; cfg=$2; shift; shift
  env "$@" timeout 100 python bench.py --config $cfg --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/fg_$name.json 2> gpurun_out/fg_$name.err
  python - gpurun_out/fg_$name.json "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],5), "staged", d.get("staged_ms_per_step"), "value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[2], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
run d2_graph d2k3n8 GSG_NO_PROFILE_EVENTS=1
run d2_eager d2k3n8 GSG_NO_PROFILE_EVENTS=1 GSG_NO_GRAPH=1
run d4_graph d4k3n7 GSG_NO_PROFILE_EVENTS=1
run d4_eager d4k3n7 GSG_NO_PROFILE_EVENTS=1 GSG_NO_GRAPH=1
run d2_default d2k3n8 A=1
timeout 120 python -m pytest tests/test_gpu_configs.py -x -q -k "config2 or config3" 2>&1 | tail -n 2
