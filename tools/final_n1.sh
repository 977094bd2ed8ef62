# round-2 evidence run on one B200: bench lines, ncu launch list of the same command, one --set full capture
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python bench.py --config recon --steps 5 --warmup 3 > gpurun_out/r2_bench_recon.json 2> gpurun_out/r2_bench_recon.err
GSG_RECON_NPTS=1e8 python bench.py --config recon --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_recon_1e8.json 2> gpurun_out/r2_bench_recon_1e8.err
python bench.py --config d4k3n7 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d4k3n7.json 2> gpurun_out/r2_bench_d4k3n7.err
python bench.py --config d2k3n8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d2k3n8.json 2> gpurun_out/r2_bench_d2k3n8.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_stream -s 9 -c 9 -o gpurun_out/r2_stream_full -f python tools/one_grad.py 6 3 8 2 > gpurun_out/r2_ncu_full.log 2>&1
for f in n1 reference_arm recon recon_1e8 d4k3n7 d2k3n8; do echo "== $f"; tail -c 600 gpurun_out/r2_bench_$f.json; echo; done
