#!/bin/bash
# mid-round check: phase cycles per pass, the segment-search tests (incl. streamed table / async prepare), stream KATs
set -u
OUT=gpurun_out
mkdir -p $OUT
bash tests/scripts/gpu_r2_prof.sh 2>&1 | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_seg.py -x -q > $OUT/r2_pytest_seg.log 2>&1; tail -5 $OUT/r2_pytest_seg.log
timeout 400 python -m pytest tests/test_gpu_full_configs.py tests/test_host_x3.py -m gpu -x -q > $OUT/r2_pytest_full.log 2>&1; tail -5 $OUT/r2_pytest_full.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_mid.json 2> $OUT/r2_bench_mid.err; python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/r2_bench_mid.json'))
    print('value', j['value'], 'e2e', j['e2e']['value'], 'plugin', j['e2e']['plugin'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2_bench_mid.err').read()[-1500:])
PY
