for pf in 1 0; do
for wl in hagen_joint_512_b8_T5 hagen_indi_64_b16_T1000; do DIFFSPLIT_B200_STREAM_RESPF=$pf python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-calls 0 > gpurun_out/r2_b18_$wl.json 2> gpurun_out/r2_b18_$wl.err; python -c "
import json
d=json.load(open(\"gpurun_out/r2_b18_$wl.json\")); print($pf, \"$wl\", d[\"precision\"], d[\"ms_per_step\"], {k:round(v[\"ms_per_step\"],3) for k,v in d[\"kernel_breakdown\"].items()})"; done
DIFFSPLIT_B200_STREAM_RESPF=$pf python tools/op_microbench.py gnconv_tf32 16 0 16 3 8 512 512
DIFFSPLIT_B200_STREAM_RESPF=$pf python tools/op_microbench.py gnconv 16 0 16 3 8 512 512
done
# evidence: launch list of the default bench command (eager launches), then the full-set capture of the persistent conv launches of one step
python bench.py --steps 2 --warmup 1 --no-graph --no-extras --no-cpu-baseline --e2e-calls 0 > gpurun_out/r2_ll_plain.json 2> gpurun_out/r2_ll_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 400 --csv --log-file gpurun_out/r2_launches_sr3_64_512_b8_eager.csv python bench.py --steps 2 --warmup 1 --no-graph --no-extras --no-cpu-baseline --e2e-calls 0 > gpurun_out/r2_ll_ncu.log 2>&1
tail -2 gpurun_out/r2_ll_ncu.log | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/r2_ll_plain.json')); print(d['ms_per_step'], d['launches_per_step'], d['gpu_launches'])"
