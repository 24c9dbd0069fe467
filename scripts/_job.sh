set -x
python -m pytest tests/test_slab_gpu.py -x -q -s 2>&1 | tail -15
