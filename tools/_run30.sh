for m in 1000 64 32; do
for wl in hagen_joint_512_b8_T5 hagen_indi_64_b16_T1000; do DIFFSPLIT_B200_UNFUSE_MINC_TF32=$m DIFFSPLIT_B200_DUMP_OPS=gpurun_out/r2_ops13_${m}_$wl.json python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b16_$wl.json 2> gpurun_out/r2_b16_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b16_$wl.json\")); print($m, \"$wl\", d[\"precision\"], d[\"ms_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "unet or seeded or psnr" 2>&1 | tail -2
