#!/bin/bash
# Segment kernel iteration: ladder (correctness), per-kind timing, phase cycles.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tests/gpu_seg_check.py small > $OUT/r2_seg_small.log 2>&1
tail -1 $OUT/r2_seg_small.log; grep MISMATCH $OUT/r2_seg_small.log | head -10
timeout 600 python tests/gpu_seg_kinds.py 6 10 > $OUT/r2_seg_kinds.log 2>&1; cat $OUT/r2_seg_kinds.log
X3_SEG_PROF=1 timeout 600 python tests/gpu_seg_kinds.py 6 10 2>&1 | grep x3_seg_kernel | awk 'NR%9>=7 || NR%9==0' > $OUT/r2_seg_phases.log; cat $OUT/r2_seg_phases.log
timeout 600 python tests/gpu_seg_check.py big > $OUT/r2_seg_big.log 2>&1; grep seg $OUT/r2_seg_big.log
if [ "${SEG_RACE:-0}" = "1" ]; then
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python tests/gpu_quick.py 60000 8192 6 nocheck C5 > $OUT/r2_seg_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|hazard" $OUT/r2_seg_racecheck.log | head -8
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tests/gpu_quick.py 60000 8192 6 check C5 > $OUT/r2_seg_memcheck.log 2>&1
grep -E "ERROR SUMMARY|oracle check" $OUT/r2_seg_memcheck.log | head -5
fi
