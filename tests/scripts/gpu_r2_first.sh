#!/bin/bash
# Round 2, first GPU call: the whole GPU suite (incl. the full-size tables and stream KATs), the new
# strong-scaling bench at N = 1, per-launch profile of C2 and C5 (before the segment kernel), and
# what page-locking a 212 MB malloc'ed buffer costs.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt

echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log

echo "== bench N=1"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r2_bench_n1_first.json 2> $OUT/r2_bench_n1_first.err
echo "bench exit $?"
cat $OUT/r2_bench_n1_first.json
tail -5 $OUT/r2_bench_n1_first.err

echo "== per-launch profile C2"
X3_RANK_PROFILE=1 X3_TRACE=1 timeout 300 python tests/gpu_quick.py 10192446 8192 5 nocheck C2 > $OUT/r2_trace_C2.log 2>&1
grep -c profile $OUT/r2_trace_C2.log

echo "== page-locking cost"
timeout 300 python - > $OUT/r2_register.log 2>&1 <<'PY'
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
pkg = g.load_package(); L = pkg.lib()
n = 211_938_580
pkg.search_host(np.zeros(100000, dtype=np.uint8))  # context up
for rep in range(3):
    a = np.empty(n + 8192, dtype=np.uint8); a[:] = 7
    t0 = time.perf_counter(); rc = L.x3s_host_register(a.ctypes.data, len(a)); t1 = time.perf_counter()
    L.x3s_host_unregister(a.ctypes.data); t2 = time.perf_counter()
    print(f"cudaHostRegister 212 MB: rc {rc} {1e3*(t1-t0):.1f} ms, unregister {1e3*(t2-t1):.1f} ms")
    t0 = time.perf_counter(); p = L.x3s_host_alloc(n); t1 = time.perf_counter(); L.x3s_host_free(p); t2 = time.perf_counter()
    print(f"cudaMallocHost 212 MB: {1e3*(t1-t0):.1f} ms, free {1e3*(t2-t1):.1f} ms")
PY
cat $OUT/r2_register.log
ls -la $OUT | tail -8
