#!/bin/bash
# compute-sanitizer racecheck / synccheck / memcheck over one case per kernel family (outputs in gpurun_out/)
OUT=gpurun_out
for tool in racecheck synccheck memcheck; do
    timeout 900 compute-sanitizer --tool $tool python tools/sanitize_kernels.py > $OUT/sanitizer_${tool}_r02.txt 2>&1
    echo "== $tool: $(grep -c ' ok$' $OUT/sanitizer_${tool}_r02.txt) cases ok; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/sanitizer_${tool}_r02.txt | tail -1)"
done
