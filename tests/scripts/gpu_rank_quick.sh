for cfg in C2 C4; do timeout 100 python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg 2>&1 | tail -1; done
timeout 100 python tests/gpu_quick.py 20000000 8192 5 nocheck C3 2>&1 | tail -1
timeout 300 python tests/gpu_rank_check.py quick 2>&1 | grep -cE " OK "
