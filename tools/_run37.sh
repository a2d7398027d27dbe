python tools/op_microbench.py conv_tc 64 0 64 3 8 512 512
MB_NO_RESIDUAL=1 python tools/op_microbench.py conv_tc 64 0 64 3 8 512 512
ncu --set full --import-source on --clock-control none -k regex:conv_tcs --launch-skip 3 -c 1 -o gpurun_out/r2_tcs_64_64_v2 python tools/op_microbench.py conv_tc 64 0 64 3 8 512 512 > gpurun_out/ncu37.log 2>&1
tail -2 gpurun_out/ncu37.log
