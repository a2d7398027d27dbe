for pf in 1 0; do
DIFFSPLIT_B200_TC_RESPF=$pf python tools/op_microbench.py conv_tc 64 0 64 3 8 512 512
DIFFSPLIT_B200_TC_RESPF=$pf python tools/op_microbench.py conv_tc 128 0 128 3 8 256 256
DIFFSPLIT_B200_TC_RESPF=$pf python tools/op_microbench.py conv_tc 256 0 256 3 8 128 128
done
