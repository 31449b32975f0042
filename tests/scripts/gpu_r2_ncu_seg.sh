#!/bin/bash
# ncu --set full of the segment kernel on one member kind (default chem) and on text; source pages kept
set -u
OUT=gpurun_out
mkdir -p $OUT
for kind in ${KINDS:-chem text}; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:x3_seg -s 1 -c 1 -f -o $OUT/r2_seg_$kind \
	python tests/gpu_one_kind.py $kind 10 3 > $OUT/r2_seg_ncu_$kind.log 2>&1
echo "ncu $kind exit $?"; tail -2 $OUT/r2_seg_ncu_$kind.log
done
ls -la $OUT/r2_seg_*.ncu-rep
