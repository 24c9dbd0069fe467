set -x
python -m pytest tests/test_replay_gpu.py tests/test_slab_gpu.py tests/test_sampler_stats_gpu.py -x -q 2>&1 | tail -5
python scripts/probe_bench.py --ept 4 --sweeps 32
