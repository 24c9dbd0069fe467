set -x
python -m pytest tests/test_replay_gpu.py -x -q 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:colour_step -s 16 -c 2 -o gpurun_out/prof_coop1 python scripts/probe_bench.py --ept 8 --sweeps 4 --warm 2 > gpurun_out/ncu_coop1.log 2>&1
tail -3 gpurun_out/ncu_coop1.log
