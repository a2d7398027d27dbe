timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_default_v2.json 2> gpurun_out/r2_bench_default_v2.err; tail -c 600 gpurun_out/r2_bench_default_v2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_default_v2.json'))
print({k:d[k] for k in ('metric','value','ms_per_step','e2e','gpu_launches','clocks','tiles_per_sec')})
print(d['roofline']); print(d['step_roofline']); print(d.get('gpu_eager_baseline')); print(d.get('cpu_baseline'))
for k,v in d.get('by_workload',{}).items(): print(k, v.get('ms_per_step'), v.get('precision'))
print(d.get('tiles'))
"
