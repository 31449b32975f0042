#!/bin/bash
# Rank search: level trace, per-launch device times (ncu launch list) for C2 and C4, new GPU tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== rank tests"
timeout 900 python -m pytest tests -m gpu -x -q -k "rank" > $OUT/pytest_rank.log 2>&1; echo "exit $?"; tail -3 $OUT/pytest_rank.log
for cfg in C2 C4; do
	echo "== trace $cfg"
	X3_TRACE=1 timeout 120 python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg 2>&1 | grep -E "level|rep 2" > $OUT/rank_trace_$cfg.log
	tail -40 $OUT/rank_trace_$cfg.log
	timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:x3_rank -c 400 --csv \
		--log-file $OUT/rank_launches_$cfg.csv python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg > /dev/null 2>&1
	echo "ncu exit $?"
done
