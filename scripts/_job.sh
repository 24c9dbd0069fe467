set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -c 3000 gpurun_out/bench_r01b.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:colour_step -s 16 -c 2 -o gpurun_out/prof_team8_lock python scripts/probe_bench.py --ept 8 --sweeps 4 --warm 2 > gpurun_out/ncu_team8_lock.log 2>&1
tail -2 gpurun_out/ncu_team8_lock.log
