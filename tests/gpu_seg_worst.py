"""Segment search on adversarial inputs (10 MB each): one repeated byte, two-symbol noise, short periods, runs --
inputs whose groups stay large through all 32 levels.  python tests/gpu_seg_worst.py [MB]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = mb << 20
rng = np.random.Generator(np.random.PCG64(7))
cases = {
    "zeros": np.zeros(n, dtype=np.uint8),
    "period7": (np.arange(n) % 7).astype(np.uint8),
    "period300": (np.arange(n) % 300 % 251).astype(np.uint8),
    "rand2": rng.integers(0, 2, n).astype(np.uint8),
    "rand4": rng.integers(0, 4, n).astype(np.uint8),
    "rand16": rng.integers(0, 16, n).astype(np.uint8),
    "rand256": rng.integers(0, 256, n).astype(np.uint8),
    "runs": np.repeat(rng.integers(0, 256, n // 64 + 1).astype(np.uint8), 64)[:n],
}
for name, data in cases.items():
    best = 1e9
    for rep in range(2):
        ls, _, tm = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_SEG, pinned=True)
        best = min(best, tm.kernel_ms)
    lr, _, tr = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK, pinned=True)
    print(f"{name:10s} {mb} MB: seg {best:9.3f} ms ({n / best / 1e3:9.1f} MB/s)  rank {tr.kernel_ms:9.3f} ms  tables equal: {bool(np.array_equal(ls, lr))}  "
          f"mean Lstar {ls.mean():.2f}", flush=True)
