#!/bin/bash
# Segment kernel, first contact: sanitizer on small inputs, the ladder against the oracle, timings.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== memcheck (small)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_quick.py 60000 8192 6 check C5 > $OUT/r2_seg_memcheck.log 2>&1
grep -E "ERROR SUMMARY|oracle check|Invalid|out of bounds" $OUT/r2_seg_memcheck.log | head -12
echo "== racecheck (small)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tests/gpu_quick.py 60000 8192 6 nocheck C5 > $OUT/r2_seg_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|hazard" $OUT/r2_seg_racecheck.log | head -12
echo "== ladder"
timeout 1200 python tests/gpu_seg_check.py small > $OUT/r2_seg_small.log 2>&1
tail -4 $OUT/r2_seg_small.log
grep -c OK $OUT/r2_seg_small.log; grep MISMATCH $OUT/r2_seg_small.log | head -20
echo "== big"
timeout 900 python tests/gpu_seg_check.py big > $OUT/r2_seg_big.log 2>&1
cat $OUT/r2_seg_big.log
