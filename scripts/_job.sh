MSKC_STATS_SCAN=1 python -m pytest tests/test_sampler_stats_gpu.py -x -q -s -k scan 2>&1 | tail -20
