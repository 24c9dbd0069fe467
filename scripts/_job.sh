set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_check.py --bx 4608 --by 4608 --rounds 3 2>&1 | grep -E "round|SEAMS|rank|mskc|Error|error" | head -20
MSKC_HALO_NCCL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/slab_check.py --bx 4608 --by 4608 --rounds 3 2>&1 | grep -E "round|SEAMS|rank|mskc|Error|error" | head -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/slab_check.py --bx 2304 --by 4608 --bz 48 --rounds 3 2>&1 | grep -E "round|SEAMS|rank|mskc|Error|error" | head -20
