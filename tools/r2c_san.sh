#!/bin/bash
mkdir -p gpurun_out
export GSG_RT_BUDGET_KB=60
timeout 110 compute-sanitizer --tool memcheck python tools/sanitize_r2.py > gpurun_out/san_memcheck.txt 2>&1; tail -n 6 gpurun_out/san_memcheck.txt
timeout 110 compute-sanitizer --tool racecheck python tools/sanitize_r2.py > gpurun_out/san_racecheck.txt 2>&1; tail -n 6 gpurun_out/san_racecheck.txt
