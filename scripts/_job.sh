set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for t in 1 4 8 16 32; do MSKC_TEAM=$t python scripts/probe_bench.py --ept 8 --sweeps 16 $( [ $t = 8 ] && echo --tiles ); done
MSKC_TEAM=8 python scripts/probe_bench.py --ept 2 4 16 --sweeps 16
MSKC_TEAM=8 python scripts/probe_bench.py --box 2304 2304 96 --ept 8 --sweeps 16
MSKC_TEAM=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:colour_step -s 16 -c 2 -o gpurun_out/prof_team8 python scripts/probe_bench.py --ept 8 --sweeps 4 --warm 2 > gpurun_out/ncu_team8.log 2>&1
