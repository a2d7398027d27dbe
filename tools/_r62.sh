DIFFSPLIT_B200_TC_PERSIST=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_tc_operator or conv_tf32" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_tc_operator or conv_tf32 or unet or seeded" 2>&1 | tail -2
MB="python tools/op_microbench.py conv_tc"
$MB 64 0 64 3 8 512 512
MB_NO_RESIDUAL=1 $MB 64 0 64 3 8 512 512
$MB 128 0 64 3 8 512 512
$MB 128 0 128 3 8 256 256
$MB 256 0 256 3 8 128 128
$MB 1024 0 1024 3 8 32 32
for wl in sr3_64_512_b8_T2000 sr3_16_128_b32_T2000 hagen_joint_512_b8_T5; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/t.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/t.json')); print('$wl', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})"; done
