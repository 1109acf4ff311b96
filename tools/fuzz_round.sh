#!/bin/bash
# Runs tools/fuzz_parity.py for a few seeds; when the REFERENCE aborts on a case (it asserts on some
# slice patterns), the case is recorded and the run resumes behind it.
#   bash tools/fuzz_round.sh "1 2 3" 60
OUT=${GRAFT_REPO_ROOT:-.}/gpurun_out
mkdir -p $OUT
for s in $1; do
    start=0
    : > $OUT/fuzz_$s.log
    for attempt in 1 2 3 4 5 6; do
        timeout 300 python tools/fuzz_parity.py --seed $s --cases 100000 --seconds $2 --start $start \
            --cursor $OUT/fuzz_cursor_$s.txt >> $OUT/fuzz_$s.log 2>&1
        rc=$?
        if [ $rc -eq 0 ] || [ $rc -eq 1 ]; then break; fi
        echo "ABORTED (rc=$rc) at: $(cat $OUT/fuzz_cursor_$s.txt)" >> $OUT/fuzz_$s.log
        start=$(( $(cut -d' ' -f1 $OUT/fuzz_cursor_$s.txt) + 1 ))
    done
    grep -c MISMATCH $OUT/fuzz_$s.log
    grep "^fuzz:\|^ABORTED" $OUT/fuzz_$s.log | cut -c1-400
done
