#!/bin/bash
# phase cycles of the segment kernel per member kind (X3_SEG_PROF=1 build of the same kernel)
set -u
OUT=gpurun_out
mkdir -p $OUT
for kind in ${KINDS:-text img chem exe}; do
X3_SEG_PROF=1 timeout 60 python tests/gpu_one_kind.py $kind 10 2 2>&1 | grep "x3_seg_kernel" | tail -2
done > $OUT/r2_seg_phases2.log
cat $OUT/r2_seg_phases2.log
