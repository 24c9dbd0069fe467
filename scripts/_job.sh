set -x
python -m pytest tests/test_replay_gpu.py -x -q 2>&1 | tail -5
for t in 8 16 32; do MSKC_TEAM=$t python scripts/probe_bench.py --ept 8 --sweeps 16; done
MSKC_TEAM=8 python scripts/probe_bench.py --ept 4 16 32 --sweeps 16
