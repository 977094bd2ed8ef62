#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/c2_parity_$i.log 2>&1; echo "rc=$?" >> gpurun_out/c2_parity_$i.log
done
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "config2 or config3 or laplacian or rk4_256 or streaming or wrappers or ode or config1 or vlasov" > gpurun_out/c2_configs.log 2>&1; echo "rc=$?" >> gpurun_out/c2_configs.log
for cfg in d2k3n8 d4k3n7; do
    timeout 300 python bench.py --config $cfg --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/c2_bench_${cfg}.json 2> gpurun_out/c2_bench_${cfg}.err
done
timeout 800 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not partition" > gpurun_out/c2_par_mc.log 2>&1
tail -n 4 gpurun_out/c2_parity_1.log gpurun_out/c2_parity_2.log gpurun_out/c2_configs.log gpurun_out/c2_par_mc.log
grep -m2 -B2 -A24 "Invalid\|out of bounds\|misaligned" gpurun_out/c2_par_mc.log | head -n 60
for f in gpurun_out/c2_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["ms_per_step"], d.get("staged_ms_per_step"), d["value"], d["gpu_launches"], d["config"].get("sweep_path"), d["roofline"] and d["roofline"]["avg_launch_ms"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
