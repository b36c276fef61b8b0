timeout 200 python -m pytest tests/test_gpu_conv_tc32.py tests/test_gpu_ops2.py -x -q -k "tc32 or thin" 2>&1 | tail -3 | cut -c1-200
for sp in 1 0; do echo "== VPS_TC32_SPLIT=$sp"; for k in fat "R50 l1 conv" Fusion "5x5 s2" conv6_1 "R50 l3"; do VPS_TC32_SPLIT=$sp timeout 60 python tools/diag_tc32.py --only "$k" 2>&1 | grep -v "^$" | cut -c30-100; done; done
