#!/bin/bash
# One GPU call that refreshes the evidence under profiles/ (run through gpurun; outputs land in gpurun_out/).
#   tools/profile_round.sh <round tag, e.g. r01>
set -u
TAG=${1:-r01}
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
cap() {  # name, kernel regex, config filter, frames
    ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $OUT/$1_$TAG \
        python tools/bench_configs.py --only "$3" --frames $4 --steps 1 > $OUT/$1_$TAG.log 2>&1
}
cap fast420 sws_fast420_rgb8 "C5 4K" 64
cap fast16  sws_fast420_rgb16 "C3 4K" 16
cap scale8  sws_scale8 "C4 8K" 16
cap generic_rgbsrc sws_generic_tile "E1 4K" 4
cap generic_upscale sws_generic_tile "X1 1080p" 4
ls -la $OUT
