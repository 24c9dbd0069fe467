set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
time python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -c 3500 gpurun_out/bench_r01c.json; tail -3 gpurun_out/bench_r01c.err
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
