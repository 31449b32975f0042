#!/bin/bash
# compute-sanitizer over the rank search (memcheck, racecheck, synccheck) on small inputs.
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
	echo "== $tool"
	timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tests/gpu_quick.py 70000 8192 5 check C1 > $OUT/sanitize_$tool.log 2>&1
	echo "exit $?"
	grep -E "ERROR SUMMARY|RACECHECK SUMMARY|oracle check|Error|hazard" $OUT/sanitize_$tool.log | head -8
done
echo "== memcheck, binary data with a long tail (tail kernel) and t=64 W=64K"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_quick.py 50000 65536 5 check C4 64 > $OUT/sanitize_memcheck2.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|oracle check" $OUT/sanitize_memcheck2.log | head
echo "== memcheck + racecheck, three lanes (chunks in flight at once, fork/join on the caller's stream)"
for tool in memcheck racecheck; do
	X3_RANK_LANES=3 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tests/gpu_quick.py 300000 8192 5 check C1 > $OUT/sanitize_lanes_$tool.log 2>&1
	echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|oracle check" $OUT/sanitize_lanes_$tool.log | head -6
done
