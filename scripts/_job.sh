set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_sampler_stats_gpu.py tests/test_stats_north_star_gpu.py tests/test_slab_gpu.py tests/test_host_gpu.py -m gpu -q -x -s -k "stats or slab or parallel or north or c2 or si_slab or la_nucleus" ) > gpurun_out/r02d_stats.log 2>&1
grep -vE "^\+" gpurun_out/r02d_stats.log | tail -60
python scripts/probe_slab.py 2>&1 | grep -v "^+"
python scripts/probe_slab.py --slab 0 2>&1 | grep -v "^+"
( timeout 900 python bench.py --steps 6 --warmup 3 ) > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d_bench_n1.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "ev/step", d["events_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["launch_ms"], d["gpu_launches"])
PY
