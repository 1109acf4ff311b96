#!/bin/bash
# scale8 / scale16 check + tuning run (one GPU call): parity tests first, then C4 / X* under the kernel's env knobs,
# then one full ncu capture of the C4 kernel.
OUT=gpurun_out
python -m pytest tests/test_scale16_gpu.py tests/test_scale8_mma_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "scale16 or scale8 or scaled or baseline or slices or tiny or fallback or range" > $OUT/s8_tests.log 2>&1
tail -15 $OUT/s8_tests.log
run() {
    echo "== $*"
    env "$@" SWS_B200_DEBUG=1 python tools/bench_configs.py --only "C4 8K" 2>&1 | grep -E "scale8:|C4" | sort -u | cut -c1-230
    env "$@" python tools/bench_configs.py --only "X1,X2,X3,X4,X5,E2" 2>&1 | grep -E "^X|^E" | cut -c1-200
}
run A=1
run SWS_B200_S8_PERSIST=1
run SWS_B200_S8_STAGES=2
ncu --set full --clock-control none --import-source on -k regex:sws_scale8 -s 3 -c 1 -f -o $OUT/scale8_r02 \
    python tools/bench_configs.py --only "C4 8K" --frames 16 --steps 1 > $OUT/scale8_r02.log 2>&1
python profiles/ncu_summarize.py $OUT/scale8_r02.ncu-rep $((7680*4320*16)) > $OUT/scale8_r02_ncu_summary.txt 2>&1
ncu -i $OUT/scale8_r02.ncu-rep --page source --csv > $OUT/scale8_r02_source.csv 2>/dev/null
rm -f $OUT/scale8_r02.ncu-rep
head -60 $OUT/scale8_r02_ncu_summary.txt
