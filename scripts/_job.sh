set -x
time timeout 900 python -m pytest tests/test_host_gpu.py -x -q 2>&1 | tail -15
