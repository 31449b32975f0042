#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, the ncu launch list of the bench
# command and one full ncu capture of the production kernel.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tests/scripts/gpu_round.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log

echo "== smoke"
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1
echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log

echo "== bench reference arm"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json

echo "== bench"
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err
cat $OUT/bench_n1.json
tail -3 $OUT/bench_n1.err

echo "== per-config kernel timing"
for v in 0 3; do for cfg in C2 C4; do
	timeout 200 python tests/gpu_quick.py 8000000 8192 $v nocheck $cfg 2>&1 | tail -1
done; done > $OUT/quick.log 2>&1
cat $OUT/quick.log

echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
	--log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
echo "launch list exit $?"

echo "== ncu --set full, production kernels (rank search) on the C2 workload: first 10 launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:x3_rank -c 10 \
	-f -o $OUT/rank_full python tests/gpu_quick.py 10192446 8192 0 nocheck C2 > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
