python -m pytest tests/test_scale8_mma_gpu.py tests/test_scale16_gpu.py -m gpu -q -x 2>&1 | tail -3
for lib in base vsw; do
  echo "== $lib"
  SWS_B200_LIB=$PWD/.ab/libswscale_b200_$lib.so python tools/bench_configs.py --only "C4 8K,X1,X2,X3,X5,E2 4K" 2>&1 | grep -E "^E|^X|^C" | cut -c1-200
done
