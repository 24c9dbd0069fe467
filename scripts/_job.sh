set -x
mkdir -p gpurun_out
exec > gpurun_out/probe_slab3.txt 2>&1
MSKC_SPLIT_TIMING=1 python scripts/probe_slab.py
MSKC_SPLIT_TIMING=1 MSKC_BTEAM=32 python scripts/probe_slab.py
MSKC_SPLIT_TIMING=1 MSKC_SKIP_HALO=1 python scripts/probe_slab.py
