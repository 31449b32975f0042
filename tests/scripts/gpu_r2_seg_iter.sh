#!/bin/bash
# Segment kernel iteration: ladder (correctness), per-kind timing, phase cycles.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python tests/gpu_seg_check.py small > $OUT/r2_seg_small.log 2>&1
tail -1 $OUT/r2_seg_small.log; grep MISMATCH $OUT/r2_seg_small.log | head -10
timeout 600 python tests/gpu_seg_kinds.py 6 10 > $OUT/r2_seg_kinds.log 2>&1; cat $OUT/r2_seg_kinds.log
X3_SEG_PROF=1 timeout 600 python tests/gpu_seg_kinds.py 6 10 2>&1 | grep x3_seg_kernel | awk 'NR%3==0' > $OUT/r2_seg_phases.log; cat $OUT/r2_seg_phases.log
timeout 600 python tests/gpu_seg_check.py big > $OUT/r2_seg_big.log 2>&1; grep seg $OUT/r2_seg_big.log
