set -x
python -m pytest tests/test_slab_gpu.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/slab_check.py --bx 4608 --by 4608 --rounds 3 2>&1 | tail -5
