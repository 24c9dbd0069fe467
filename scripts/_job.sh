set -x
mkdir -p gpurun_out
for n in 8 4 2 1; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n"
  ( timeout 600 $TR bench.py --gpus $n --steps 6 --warmup 3 ) > gpurun_out/r02h_bench_n$n.json 2> gpurun_out/r02h_bench_n$n.err
  tail -c 200 gpurun_out/r02h_bench_n$n.err
done
