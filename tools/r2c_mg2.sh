#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mg2_$name.json 2> gpurun_out/mg2_$name.err
  python - gpurun_out/mg2_$name.json "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", round(d["ms_per_step"],4), "value", d["value"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "ERR", e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
run rowtile A=1
run norowtile GSG_ROWTILE=0
