set -x
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_check.py --bx 1152 --by 1152 --rounds 2 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/slab_check.py --bx 4608 --by 4608 --rounds 3 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 scripts/slab_check.py --bx 4608 --by 4608 --rounds 3 2>&1 | tail -8
