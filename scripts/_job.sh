set -x
mkdir -p gpurun_out
: > gpurun_out/ab_team_regs.txt
for lib in "" scripts/_build/libmskc_minb6.so scripts/_build/libmskc_minb8.so; do
  for team in 8 16 32; do
    echo "lib=${lib:-default(minb4)} team=$team" >> gpurun_out/ab_team_regs.txt
    MSKC_LIB=${lib:+$PWD/$lib} MSKC_TEAM=$team timeout 300 python scripts/probe_bench.py --ept 4 --sweeps 16 --warm 8 >> gpurun_out/ab_team_regs.txt 2>&1
  done
done
cat gpurun_out/ab_team_regs.txt
for c in surface CVD-Si-SiH2; do python tests/probe_parity.py $c 400000; done 2>&1 | tee gpurun_out/r02_parity_rate.txt
( time python -m pytest tests/test_stats_north_star_gpu.py -m gpu -q -s ) > gpurun_out/stats_north_star.log 2>&1
grep -vE "^\+" gpurun_out/stats_north_star.log | tail -45
