python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; tail -c 300 gpurun_out/r2_bench_default_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm_final.json 2> gpurun_out/r2_bench_reference_arm_final.err; tail -c 300 gpurun_out/r2_bench_reference_arm_final.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_default_final.json'))
print({k:d[k] for k in ('metric','value','ms_per_step','e2e','gpu_launches','clocks','tiles_per_sec','config')})
print(d['roofline']); print(d['step_roofline']); print(d.get('gpu_eager_baseline')); print(d.get('cpu_baseline'))
for k,v in d.get('by_workload',{}).items(): print(k, v.get('ms_per_step'), v.get('precision'))
print(d.get('tiles'))
r=json.load(open('gpurun_out/r2_bench_reference_arm_final.json')); print(r)
"
