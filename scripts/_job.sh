set -x
mkdir -p gpurun_out
python scripts/probe_m.py 2>&1 | grep "^box"
( timeout 1500 python -m pytest tests/test_stats_north_star_gpu.py tests/test_sampler_stats_gpu.py tests/test_slab_gpu.py -m gpu -q -s ) > gpurun_out/r02g_stats.log 2>&1
grep -vE "^\+" gpurun_out/r02g_stats.log | grep -E "^[A-Z]|  [a-z_A-Z]+ |passed|failed" | tail -45
python scripts/probe_slab.py --slab 0 2>&1 | grep -v "^+"
