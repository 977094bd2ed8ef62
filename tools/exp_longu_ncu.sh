GSG_LONGU_PMAX=6 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_longu_list.csv python tools/one_grad.py 6 3 8 2 > /dev/null 2>&1
GSG_NO_LONGU=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_nolongu_list.csv python tools/one_grad.py 6 3 8 2 > /dev/null 2>&1
GSG_LONGU_PMAX=6 ncu --set full --clock-control none --import-source on -k regex:sweep_longu -s 4 -c 4 -o gpurun_out/prof_longu -f python tools/one_grad.py 6 3 8 2 > gpurun_out/ncu_longu_full.log 2>&1
python - <<'PY'
import csv, collections
for name in ("longu", "nolongu"):
    rows = [r for r in csv.reader(open(f"gpurun_out/ncu_{name}_list.csv")) if len(r) > 10 and r[0].isdigit()]
    half = rows[len(rows)//2:]
    agg = collections.OrderedDict()
    for r in half:
        k = (r[4].split("(")[0][-40:], r[8], r[7])
        agg.setdefault(k, []).append(float(r[-1]) / 1e3)
    print("==", name, "second RHS: kernel, grid, block -> n, avg us, total us")
    tot = 0
    for k, v in agg.items():
        print("  ", k, len(v), "%.1f" % (sum(v)/len(v)), "%.1f" % sum(v)); tot += sum(v)
    print("   total %.1f us" % tot)
PY
