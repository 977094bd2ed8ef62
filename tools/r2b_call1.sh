#!/bin/bash
# round 2, session 2, GPU call 1: parity of the flat path + pre-squared Laplacian, small-config bench lines (flat vs
# tiled), sanitizer on the flat kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/c1_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/c1_parity.log
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "config2 or config3 or laplacian or rk4_256 or streaming or wrappers or ode or config1 or vlasov" > gpurun_out/c1_configs.log 2>&1
echo "configs rc=$?" >> gpurun_out/c1_configs.log
for cfg in d2k3n8 d4k3n7; do
  for f in 0 2; do
    GSG_FLAT=$f timeout 300 python bench.py --config $cfg --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_${cfg}_flat$f.json 2> gpurun_out/c1_bench_${cfg}_flat$f.err
  done
done
GSG_SAN_MODES=1 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/c1_memcheck_flat.log 2>&1
GSG_SAN_MODES=1 timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/c1_racecheck_flat.log 2>&1
tail -3 gpurun_out/c1_parity.log gpurun_out/c1_configs.log
tail -2 gpurun_out/c1_memcheck_flat.log gpurun_out/c1_racecheck_flat.log
for f in gpurun_out/c1_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["ms_per_step"], d.get("staged_ms_per_step"), d["value"], d["gpu_launches"], d["config"].get("sweep_path"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
