set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:parity_coop -s 1 -c 1 -o gpurun_out/parity_coop -f python scripts/probe_parity.py surface 30000 > gpurun_out/ncu_parity.log 2>&1
tail -3 gpurun_out/ncu_parity.log
