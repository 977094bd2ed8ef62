run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp3_$name.json 2> gpurun_out/exp3_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/exp3_$name.json").read().strip().splitlines()[-1]); print("$name", "%.3f ms" % d["ms_per_step"], "staged %.3f" % d["staged_ms_per_step"], "stream avg %.1f us" % (1e3*d["roofline"]["avg_launch_ms"]), "launches", d["gpu_launches"], "e2e %.3e" % d["e2e"]["value"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/exp3_$name.err").read()[-600:])
PY
}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "fused_gradient or config4 or mg_ or rk4 or ode" 2>&1 | grep -E "passed|failed|FAILED" | head
run conc2 A=1
run conc2_onestream GSG_RHS_ONE_STREAM=1
run serial GSG_RHS_SERIAL=1
