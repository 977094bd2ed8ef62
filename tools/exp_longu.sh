run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp2_$name.json 2> gpurun_out/exp2_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/exp2_$name.json").read().strip().splitlines()[-1]); print("$name", "%.3f ms" % d["ms_per_step"], "staged %.3f" % d["staged_ms_per_step"], "stream avg %.1f us" % (1e3*d["roofline"]["avg_launch_ms"]), "launches", d["gpu_launches"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/exp2_$name.err").read()[-600:])
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -k "apply_D_every_axis or fused_gradient or config4 or config2_apply or config3_apply or streaming_kernel" 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -20
run pmax5 GSG_LONGU_PMAX=5
run pmax6 A=1
run pmax5_g2 GSG_LONGU_PMAX=5 GSG_LONGU_GRID=2
run pmax5_g1 GSG_LONGU_PMAX=5 GSG_LONGU_GRID=1
run pmin5_pmax5 GSG_LONGU_PMIN=5 GSG_LONGU_PMAX=5
run pmax4 GSG_LONGU_PMAX=4
GSG_LONGU_PMAX=6 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_longu_list.csv python tools/one_grad.py 6 3 8 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for name in ("longu",):
    rows = [r for r in csv.reader(open(f"gpurun_out/ncu_{name}_list.csv")) if len(r) > 10 and r[0].isdigit()]
    half = rows[len(rows)//2:]
    agg = collections.OrderedDict()
    for r in half:
        k = (r[4].split("(")[0][-40:], r[8], r[7])
        agg.setdefault(k, []).append(float(r[-1]) / 1e3)
    print("==", name, "second RHS: kernel, grid, block -> n, avg us, total us")
    for k, v in agg.items():
        print("  ", k, len(v), "%.1f" % (sum(v)/len(v)), "%.1f" % sum(v))
PY
