set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r01b_pytest_gpu.log 2>&1
tail -5 gpurun_out/r01b_pytest_gpu.log
python bench.py > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err
tail -c 3000 gpurun_out/r01b_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01b_bench_under_ncu.log 2>&1
wc -l gpurun_out/r01b_launches.csv
