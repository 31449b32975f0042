#!/bin/bash
# Round 2, one gpurun call: GPU suite, smoke, both bench arms at N = 1, the ncu launch list of the
# bench command and one --set full capture of the production launch (segment kernel on all of C5).
#   gpurun --timeout 1700 -- 'bash tests/scripts/gpu_r2_round.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt

echo "== pytest -m gpu"
timeout 700 python -m pytest tests -m gpu -x -q --durations=10 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -18 $OUT/pytest_gpu.log

echo "== smoke"
timeout 200 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log

echo "== bench reference arm"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2_bench_ref.json 2> $OUT/r2_bench_ref.err
cat $OUT/r2_bench_ref.json

echo "== bench N=1"
timeout 600 python bench.py > $OUT/r2_bench_n1.json 2> $OUT/r2_bench_n1.err
echo "bench exit $?"
cat $OUT/r2_bench_n1.json
tail -5 $OUT/r2_bench_n1.err

echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
	--log-file $OUT/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/r2_launches_bench.log 2>&1
echo "launch list exit $?"

echo "== ncu --set full: the segment kernel on all of C5, device resident (the launch the bench's device leg times)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:x3_seg -s 1 -c 1 \
	-f -o $OUT/r2_seg_c5 python tests/gpu_c5_device.py C5 2 > $OUT/r2_seg_ncu_c5.log 2>&1
echo "ncu full exit $?"; tail -3 $OUT/r2_seg_ncu_c5.log
ls -la $OUT | tail -12
