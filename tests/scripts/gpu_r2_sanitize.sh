#!/bin/bash
# compute-sanitizer over the segment search (racecheck, memcheck, synccheck)
set -u
OUT=gpurun_out
mkdir -p $OUT
for tool in racecheck memcheck synccheck; do
timeout 500 compute-sanitizer --tool $tool --print-limit 10 python tests/gpu_seg_sanitize.py > $OUT/r2_seg_$tool.log 2>&1
echo "$tool exit $?"; grep -E "SUMMARY|oracle check|pieces in turn|hazard|Error" $OUT/r2_seg_$tool.log | head -12
done
