for lag in 1 2 3 4; do echo "lag $lag"; for cfg in C2 C4; do X3_RANK_LAG=$lag timeout 100 python tests/gpu_quick.py 10192446 8192 5 nocheck $cfg 2>&1 | tail -1 | cut -c1-120; done; done
