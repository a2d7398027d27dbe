timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
DIFFSPLIT_B200_TC_PERSIST=2 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "conv_tc_operator or conv_tf32 or unet or seeded" 2>&1 | tail -2
DIFFSPLIT_B200_NO_RES_FOLD=1 DIFFSPLIT_B200_NO_ENTRY_TC=1 DIFFSPLIT_B200_TC_RESPF=0 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "unet or seeded" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "full_resolution" 2>&1 | grep "full resolution" 
