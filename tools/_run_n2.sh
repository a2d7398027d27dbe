timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "another_gpu or packed" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
tail -c 400 gpurun_out/r2_bench_n2_final.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2_final.json'))
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','tiles_per_sec','scaling')})
print(d.get('tiles'))
"
