for wl in hagen_joint_512_b8_T5 hagen_indi_64_b16_T1000; do DIFFSPLIT_B200_DUMP_OPS=gpurun_out/r2_ops12_$wl.json python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b15_$wl.json 2> gpurun_out/r2_b15_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b15_$wl.json\")); print(\"$wl\", d[\"precision\"], d[\"ms_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done
