set -x
mkdir -p gpurun_out
exec > gpurun_out/ab_loads.txt 2>&1
( timeout 600 python -m pytest tests/test_slab_gpu.py tests/test_host_gpu.py -m gpu -q -x -k "slab or parallel" ) 2>&1 | tail -2
for lib in "" scripts/_build/libmskc_prefetch.so scripts/_build/libmskc_early.so scripts/_build/libmskc_early_pf.so; do
  echo "== lib=${lib:-default}"
  MSKC_LIB=${lib:+$PWD/$lib} python scripts/probe_slab.py --slab 0 --sweeps 64
done
echo "== default again"
python scripts/probe_slab.py --slab 0 --sweeps 64
