set -x
(time python -m pytest tests -x -q -m gpu 2>&1 | tail -5) 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -c 3000 gpurun_out/bench_r01d.json; tail -3 gpurun_out/bench_r01d.err
