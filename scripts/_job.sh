set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02d_pytest_gpu.log 2>&1
tail -12 gpurun_out/r02d_pytest_gpu.log
( timeout 1200 python -m pytest tests/test_stats_north_star_gpu.py tests/test_sampler_stats_gpu.py -m gpu -q -s ) > gpurun_out/r02d_stats.log 2>&1
grep -vE "^\+" gpurun_out/r02d_stats.log | grep -E "^[A-Z]|  [a-z_A-Z]+ |passed|failed" | tail -40
