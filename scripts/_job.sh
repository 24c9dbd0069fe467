# what was run on the GPU box (gpurun -- 'bash scripts/_job.sh')
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python bench.py --no-la-cycle > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python scripts/probe_bench.py --ept 4 --sweeps 16 --warm 24 --tiles
ncu --set full --clock-control none --import-source on -k regex:colour_step -s 80 -c 2 -o gpurun_out/colour_step -f python scripts/probe_bench.py --ept 4 --sweeps 8 --warm 24 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
