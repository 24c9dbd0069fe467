set -x
P="python scripts/probe_bench.py --ept 4 --sweeps 32 --warm 16 --box 9216 4608 96"
for pad in 0 20000 30000 70000; do MSKC_SMEM_PAD=$pad $P 2>&1 | tail -1; done
