set -x
time MSKC_FULL_GOLDEN=1 python -m pytest tests/test_host_gpu.py -x -q -k "cube-zigzag or sphere" 2>&1 | tail -5
