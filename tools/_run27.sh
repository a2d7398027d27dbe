DIFFSPLIT_B200_TC_PERSIST=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "conv_tc_operator or conv_tf32" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "conv_tc_operator or conv_tf32 or unet or seeded" 2>&1 | tail -2
MB="python tools/op_microbench.py conv_tc"
for pf in 1 0; do
DIFFSPLIT_B200_TC_RESPF=$pf $MB 64 0 64 3 8 512 512
DIFFSPLIT_B200_TC_RESPF=$pf $MB 128 0 128 3 8 256 256
DIFFSPLIT_B200_TC_RESPF=$pf $MB 256 0 256 3 8 128 128
DIFFSPLIT_B200_TC_RESPF=$pf $MB 128 0 64 3 8 512 512
done
MB_NO_RESIDUAL=1 $MB 64 0 64 3 8 512 512
MB_NO_RESIDUAL=1 $MB 128 0 128 3 8 256 256
for wl in sr3_64_512_b8_T2000 sr3_16_128_b32_T2000; do DIFFSPLIT_B200_DUMP_OPS=gpurun_out/r2_ops11_$wl.json python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b14_$wl.json 2> gpurun_out/r2_b14_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b14_$wl.json\")); print(\"$wl\", d[\"precision\"], d[\"ms_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done
