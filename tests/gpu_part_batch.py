"""x3s_search_host_part (part 0 of N) on C5 from page-locked buffers under X3_PART_BATCH (pieces per launch).
    python tests/gpu_part_batch.py"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

corpus = g.load_submodule("corpus")
data = np.frombuffer(corpus.generate_cached("C5"), dtype=np.uint8)
pkg = g.load_package()
L = pkg.lib()
n, W, t = len(data), 8192, 15
px, pl = L.x3s_host_alloc(n + W), L.x3s_host_alloc(n)
x = np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_uint8)), shape=(n + W,))
x[:n] = data
x[n:] = 0
tm = pkg.Timing()
for parts in (2, 8):
    for per in ("default", 2, 4, 8, 13, 26, 53):
        if per == "default":
            os.environ.pop("X3_PART_BATCH", None)
        else:
            os.environ["X3_PART_BATCH"] = str(per)
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            assert L.x3s_search_host_part(px, n, W, t, pl, C.byref(tm), 0, parts) == 0
            best = min(best, (time.perf_counter() - t0) * 1e3)
        print(f"parts {parts} batch {per}: {best:.3f} ms, {tm.launches} launches, kernel part {tm.kernel_ms:.3f} ms", flush=True)
