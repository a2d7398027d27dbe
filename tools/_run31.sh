for pe in 0 2; do DIFFSPLIT_B200_TC_PATCH=$pe DIFFSPLIT_B200_TC_PERSIST=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "conv_tc_operator or conv_tf32" 2>&1 | tail -2; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "conv_tc_operator or conv_tf32 or unet or seeded" 2>&1 | tail -2
for pm in 148 64 24; do
for wl in hagen_joint_512_b8_T5 hagen_indi_64_b16_T1000 sr3_16_128_b32_T2000 sr3_64_512_b8_T2000; do DIFFSPLIT_B200_TC_PERSIST_MIN=$pm DIFFSPLIT_B200_DUMP_OPS=gpurun_out/r2_ops14_${pm}_$wl.json python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b17_$wl.json 2> gpurun_out/r2_b17_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b17_$wl.json\")); print($pm, \"$wl\", d[\"precision\"], d[\"ms_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done; done
