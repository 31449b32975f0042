#!/bin/bash
# A/B of two builds of the library: lib/ (A) and lib_ab/ (B, copied over A on the box's scratch copy).
set -u
run() {
	for cfg in C2 C4; do timeout 100 python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg 2>&1 | tail -1; done
	timeout 100 python tests/gpu_quick.py 20000000 8192 5 nocheck C3 2>&1 | tail -1
	timeout 100 python tests/gpu_quick.py 40000000 8192 5 nocheck C5 2>&1 | tail -1
}
echo "== A"; run
cp x3-compressor_b200/lib_ab/libx3b200.so x3-compressor_b200/lib/libx3b200.so
echo "== B"; run
timeout 300 python tests/gpu_rank_check.py quick 2>&1 | grep -E "MISMATCH|ALL OK|FAILED" | head
