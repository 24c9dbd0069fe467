set -x
for t in 4 3 2; do MSKC_TPW=$t python scripts/probe_bench.py --ept 4 --sweeps 32 --warm 16 2>&1 | tail -1; done
