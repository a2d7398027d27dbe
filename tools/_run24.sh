for minc in 64 128; do
for wl in sr3_64_512_b8_T2000 sr3_16_128_b32_T2000; do DIFFSPLIT_B200_UNFUSE_MINC=$minc DIFFSPLIT_B200_DUMP_OPS=gpurun_out/r2_ops10_${minc}_$wl.json python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b13_${minc}_$wl.json 2> gpurun_out/r2_b13_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b13_${minc}_$wl.json\")); print($minc, \"$wl\", d[\"precision\"], d[\"ms_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done; done
