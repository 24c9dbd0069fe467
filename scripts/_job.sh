set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:colour_step -s 120 -c 1 -o gpurun_out/prof_r01e python scripts/probe_bench.py --ept 4 --sweeps 32 --warm 16 > gpurun_out/ncu_r01e.log 2>&1
tail -3 gpurun_out/ncu_r01e.log
