TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; n=$2; shift; shift; env "$@" timeout 240 $TR --nproc-per-node $n --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/mg_$name.json 2> gpurun_out/mg_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/mg_$name.json").read().strip().splitlines()[-1]); print("$name", "%.3f ms" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "stream avg %.1f us" % (1e3*d["roofline"]["avg_launch_ms"]) if d.get("roofline") else "")
except Exception as e: print("$name failed", e); print(open("gpurun_out/mg_$name.err").read()[-800:])
PY
}
run n8 8 A=1
run n8_rs8 8 GSG_LONG_RSPLIT=8
run n8_nozero 8 GSG_MG_NO_ZERO=1
run n4 4 A=1
run n2 2 A=1
