#!/bin/bash
# Full ncu captures with the source page, summarised on the box (reports are too big to bring back):
#   tools/ncu_capture.sh <tag> "<name> <kernel regex> <config filter> <frames> <pixels per launch>" ...
TAG=$1; shift
OUT=gpurun_out
for spec in "$@"; do
    set -- $spec
    name=$1; kre=$2; cfg=$3; frames=$4; px=$5
    ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -f -o $OUT/${name}_$TAG \
        python tools/bench_configs.py --only "$cfg" --frames $frames --steps 1 > $OUT/${name}_$TAG.log 2>&1
    python profiles/ncu_summarize.py $OUT/${name}_$TAG.ncu-rep $px > $OUT/${name}_${TAG}_ncu_summary.txt 2>&1
    ncu -i $OUT/${name}_$TAG.ncu-rep --page source --csv > $OUT/${name}_${TAG}_source.csv 2>/dev/null
    rm -f $OUT/${name}_$TAG.ncu-rep
    head -22 $OUT/${name}_${TAG}_ncu_summary.txt
done
