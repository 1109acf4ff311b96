#!/bin/bash
# One GPU call that refreshes the evidence under profiles/ (run through gpurun; outputs land in gpurun_out/).
#   tools/profile_round.sh <round tag, e.g. r01>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python tools/bench_configs.py > $OUT/configs_$TAG.txt 2>&1
python tools/bench_configs.py --frames 16 > $OUT/configs_16frames_$TAG.txt 2>&1
python bench.py --steps 20 --warmup 3 > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference_arm.json 2>> $OUT/bench_${TAG}_n1.err
# launch list of the bench command (per-launch durations are cold/serialised: only the SHARE matters)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# full captures of the dominant kernel of every BASELINE configuration
cap() {  # name, kernel regex, config filter, frames, pixels per launch (thread-instructions per pixel in the summary)
    ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/$1_$TAG \
        python tools/bench_configs.py --only "$3" --frames $4 --steps 1 > $OUT/$1_$TAG.log 2>&1
    # the reports are 10+ MB each and gpurun brings back at most 64 MiB: summarise here, keep the text
    python profiles/ncu_summarize.py $OUT/$1_$TAG.ncu-rep $5 > $OUT/$1_${TAG}_ncu_summary.txt 2>&1
    rm -f $OUT/$1_$TAG.ncu-rep
}
cap fast420 sws_fast420_rgb8 "C5 4K" 64 $((3840*2160*64))
cap fast16  sws_fast420_rgb16 "C3 4K" 16 $((3840*2160*16))
cap scale8  sws_scale8 "C4 8K" 16 $((7680*4320*16))
cap scale8_rgb sws_scale8 "X1 1080p" 16 $((3840*2160*16))
cap rgb420 sws_rgb420 "E1 4K" 16 $((3840*2160*16))
cap fast_hi8 sws_fast420_hi8 "C3b 4K" 16 $((3840*2160*16))
cap scale_rgb sws_scale8 "E2 4K" 16 $((3840*2160*16))
cap scale16 sws_scale8 "X3 4K" 16 $((3840*2160*16))
cap scale8_x2 sws_scale8 "X2 4K" 16 $((3840*2160*16))
cap scale_i19 sws_scale8 "X6 4K" 16 $((3840*2160*16))
cap scale_p010 sws_scale8 "X8 4K" 16 $((3840*2160*16))
cap scale_rgb48 sws_scale8 "X10 4K" 16 $((3840*2160*16))
cap copy8 sws_copy8 "U1 4K" 16 $((3840*2160*16))
cap full444 sws_full444 "F1 4K" 16 $((3840*2160*16))
cap rgb444 sws_rgb444 "F3 4K" 16 $((3840*2160*16))
ls -la $OUT
