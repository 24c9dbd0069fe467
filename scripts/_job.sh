set -x
mkdir -p gpurun_out
python scripts/probe_slab.py 2>&1 | grep -v "^+"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/r02_slab_launches.csv python scripts/probe_slab.py > gpurun_out/r02_slab_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02_slab_launches.csv', errors='ignore')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; col = {n: i for i, n in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[col['Metric Name']] != 'gpu__time_duration.sum': continue
    k = r[col['Kernel Name']][:70]; v = float(r[col['Metric Value']].replace(',', '')); u = r[col['Metric Unit']]
    v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]): print(f"{a[0]:6d} {a[1]:10.1f} us  avg {a[1]/a[0]:8.2f}  {k}")
PY
( timeout 900 python bench.py ) > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02e_bench_n1.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "ev/step", d["events_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["launch_ms"], d["roofline"]["share_of_step"], d["gpu_launches"])
PY
