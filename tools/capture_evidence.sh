# Evidence captures of profiles/ (round 2): run on the GPU box as  gpurun -- bash tools/capture_evidence.sh ; outputs in gpurun_out/
set -x
B="python bench.py --steps 2 --warmup 1 --no-graph --no-extras --no-cpu-baseline --e2e-calls 0"
$B > gpurun_out/r2_ev_plain.json 2> gpurun_out/r2_ev_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 400 --csv --log-file gpurun_out/r2_launches_sr3_64_512_b8_eager_v2.csv $B > gpurun_out/r2_ev_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed -k regex:conv_tc --clock-control none --launch-skip 45 -c 90 --csv --log-file gpurun_out/r2_conv_tc_dram_per_launch.csv $B > gpurun_out/r2_ev_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv_tcs --launch-skip 50 -c 6 -o gpurun_out/r2_ncu_full_conv_tcs $B > gpurun_out/r2_ev_ncu3.log 2>&1
tail -2 gpurun_out/r2_ev_ncu3.log
ls -la gpurun_out/r2_ncu_full_conv_tcs.ncu-rep
