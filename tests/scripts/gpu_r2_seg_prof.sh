#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== per kind, seg"
timeout 600 python tests/gpu_seg_kinds.py 6 10 > $OUT/r2_seg_kinds.log 2>&1; cat $OUT/r2_seg_kinds.log
echo "== per kind, rank"
timeout 600 python tests/gpu_seg_kinds.py 5 10 > $OUT/r2_rank_kinds.log 2>&1; cat $OUT/r2_rank_kinds.log
echo "== ncu full, seg kernel on C2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:x3_seg -s 1 -c 1 -f -o $OUT/r2_seg_c2 \
	python tests/gpu_quick.py 10192446 8192 6 nocheck C2 > $OUT/r2_seg_ncu_c2.log 2>&1
echo "ncu exit $?"
echo "== ncu full, seg kernel on chem 10 MB"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:x3_seg -s 1 -c 1 -f -o $OUT/r2_seg_chem \
	python tests/gpu_seg_kinds.py 6 10 > $OUT/r2_seg_ncu_chem.log 2>&1
echo "ncu exit $?"
ls -la $OUT/*.ncu-rep
