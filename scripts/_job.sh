set -x
python -m pytest tests/test_parity_gpu.py tests/test_replay_gpu.py -x -q -k "laser" 2>&1 | tail -15
