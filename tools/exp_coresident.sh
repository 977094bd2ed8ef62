run() { name=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp1_$name.json 2> gpurun_out/exp1_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/exp1_$name.json").read().strip().splitlines()[-1]); print("$name", "%.3f ms" % d["ms_per_step"], "staged %.3f" % d["staged_ms_per_step"], "stream avg %.1f us" % (1e3*d["roofline"]["avg_launch_ms"]))
except Exception as e: print("$name failed", e)
PY
}
run base A=1
run ns3 GSG_STREAM_NS=3
run ns3_c2_x64_nw4 GSG_STREAM_NS=3 GSG_LONG_C=2 GSG_LONG_XCAP=64 GSG_LONG_PASSES=1 GSG_LONG_NW=4
run ns3_c2_x64_nw8 GSG_STREAM_NS=3 GSG_LONG_C=2 GSG_LONG_XCAP=64 GSG_LONG_PASSES=1 GSG_LONG_NW=8
run ns3_c1_x64_nw4 GSG_STREAM_NS=3 GSG_LONG_C=1 GSG_LONG_XCAP=64 GSG_LONG_PASSES=1 GSG_LONG_NW=4
run ns3_c2_x64_nw4_noprio GSG_STREAM_NS=3 GSG_LONG_C=2 GSG_LONG_XCAP=64 GSG_LONG_PASSES=1 GSG_LONG_NW=4 GSG_NO_PRIO=1
run ns3_c2_x64_nw4_sfirst GSG_STREAM_NS=3 GSG_LONG_C=2 GSG_LONG_XCAP=64 GSG_LONG_PASSES=1 GSG_LONG_NW=4 GSG_STREAM_FIRST=1
run ns4_c2_x32_nw4 GSG_LONG_C=2 GSG_LONG_XCAP=32 GSG_LONG_PASSES=1 GSG_LONG_NW=4
run ns3_c2_x64_nw4_rs GSG_STREAM_NS=3 GSG_LONG_C=2 GSG_LONG_XCAP=64 GSG_LONG_PASSES=1 GSG_LONG_NW=4 GSG_LONG_RSPLIT=8
