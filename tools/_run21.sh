MB="python tools/op_microbench.py conv_tc"
$MB 128 0 64 3 8 512 512
MB_NO_RESIDUAL=1 $MB 128 0 64 3 8 512 512
MB_NO_RESIDUAL=1 MB_NO_F32=1 $MB 128 0 64 3 8 512 512
MB_NO_F32=1 $MB 128 0 64 3 8 512 512
$MB 128 0 128 3 8 256 256
MB_NO_RESIDUAL=1 MB_NO_F32=1 $MB 128 0 128 3 8 256 256
$MB 256 0 256 3 8 128 128
MB_NO_RESIDUAL=1 MB_NO_F32=1 $MB 256 0 256 3 8 128 128
$MB 1024 0 1024 3 8 32 32
MB_NO_RESIDUAL=1 MB_NO_F32=1 $MB 1024 0 1024 3 8 32 32
$MB 192 0 64 3 8 512 512
ncu --set full --import-source on --clock-control none -k regex:conv_tcs --launch-skip 3 -c 1 -o gpurun_out/r2_tcs_128_64 python tools/op_microbench.py conv_tc 128 0 64 3 8 512 512 > gpurun_out/ncu21.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv_tcs --launch-skip 3 -c 1 -o gpurun_out/r2_tcs_256_256 python tools/op_microbench.py conv_tc 256 0 256 3 8 128 128 >> gpurun_out/ncu21.log 2>&1
tail -3 gpurun_out/ncu21.log
