set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
python -c "
import json
d = json.load(open('gpurun_out/bench_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'la', d['e2e_la_cycle'])
print('roofline', d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline']['events_per_launch'])
for p in d['dense_passes']: print(p['kernel'], round(p['launch_ms'],3), 'ms', round(p['achieved']), 'GB/s', round(p['frac'],3))
print('strong', d['strong_1e9'])
print('cpu', d['cpu_baseline'])
"
