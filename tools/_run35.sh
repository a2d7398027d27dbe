timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_tf32 or unet or seeded or psnr" 2>&1 | tail -3
for ne in 0 1; do
for wl in sr3_64_512_b8_T2000 sr3_16_128_b32_T2000 hagen_joint_512_b8_T5; do if [ $ne = 1 ]; then export DIFFSPLIT_B200_NO_ENTRY_TC=1; else unset DIFFSPLIT_B200_NO_ENTRY_TC; fi; DIFFSPLIT_B200_DUMP_OPS=gpurun_out/r2_ops16_${ne}_$wl.json python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b20_$wl.json 2> gpurun_out/r2_b20_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b20_$wl.json\")); print($ne, \"$wl\", d[\"precision\"], d[\"ms_per_step\"], d[\"launches_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done; done
