set -x
mkdir -p gpurun_out
python scripts/probe_parity.py surface 200000
python scripts/probe_parity.py Si-SL 200000
python scripts/probe_parity.py CVD-Si-SiH2 200000
( python -m pytest tests/test_host_gpu.py tests/test_parity_gpu.py -m gpu -x -q ) 2>&1 | tail -3
