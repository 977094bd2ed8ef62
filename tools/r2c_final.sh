#!/bin/bash
# round-2 final evidence on one B200: full GPU suite, bench lines, ncu launch list of the bench command, one --set full
# capture of the row-tile kernel
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 > gpurun_out/f_pytest.txt; cat gpurun_out/f_pytest.txt
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 900 gpurun_out/r2_bench_n1.json
timeout 100 python bench.py --config d4k3n7 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d4k3n7.json 2> gpurun_out/r2_bench_d4k3n7.err
timeout 100 python bench.py --config d2k3n8 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d2k3n8.json 2> gpurun_out/r2_bench_d2k3n8.err
for f in d4k3n7 d2k3n8; do python - gpurun_out/r2_bench_$f.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d.get("staged_ms_per_step"), d["value"], d["gpu_launches"])
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 140 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:sweep_rowtile -s 1 -c 1 -f -o gpurun_out/r2_rowtile_full python tools/one_apply.py 6 3 8 1 2 > gpurun_out/f_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_rowtile_full.ncu-rep > gpurun_out/r2_rowtile_ncu_summary.txt 2>&1; head -30 gpurun_out/r2_rowtile_ncu_summary.txt
