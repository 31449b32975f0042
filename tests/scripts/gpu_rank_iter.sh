#!/bin/bash
# Rank search iteration: correctness sweep + per-launch device times for C2 and C4.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tests/gpu_rank_check.py quick > $OUT/rank_check.log 2>&1; echo "check exit $?"
grep -c " OK " $OUT/rank_check.log; grep -E "MISMATCH|first bad|Error|error|stream kernel|vs stream|ALL OK|FAILED" $OUT/rank_check.log | head -30
for cfg in C2 C4; do
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:x3_rank -c 400 --csv \
		--log-file $OUT/rank_launches_$cfg.csv python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg > $OUT/rank_quick_$cfg.log 2>&1
	echo "ncu exit $?"
	timeout 100 python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg 2>&1 | tail -1
done
