set -x
mkdir -p gpurun_out
python scripts/probe_bench.py --ept 4 --sweeps 16 --warm 24 2>&1 | tail -1
python scripts/probe_bench.py --ept 4 --sweeps 16 --warm 24 2>&1 | tail -1
( python -m pytest tests/test_replay_gpu.py tests/test_sampler_stats_gpu.py -m gpu -x -q ) 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:colour_step -s 80 -c 2 -o gpurun_out/colour_step3 -f python scripts/probe_bench.py --ept 4 --sweeps 8 --warm 24 > gpurun_out/ncu_full.log 2>&1
