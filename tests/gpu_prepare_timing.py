"""x3_search_prepare + x3_search_wait on a malloc'ed C5 (the plug-in call), repeated, under the table knobs.
    python tests/gpu_prepare_timing.py [reps]"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
corpus = g.load_submodule("corpus")
data = np.frombuffer(corpus.generate_cached("C5"), dtype=np.uint8)
pkg = g.load_package()
L = pkg.lib()
n, W = len(data), 8192
libc = C.CDLL(None)
libc.malloc.restype = C.c_void_p
libc.malloc.argtypes = [C.c_size_t]
buf = libc.malloc(n + W)
C.memmove(buf, data.ctypes.data, n)
C.memset(buf + n, 0, W)
L.set_forward_window(W); L.set_max_match_count(15)
for knobs in ({}, {"X3_TABLE_NO_THP": "1"}, {"X3_TABLE_KEEP": "1"}, {"X3_PREPARE_SYNC": "1"}):
    for k in ("X3_TABLE_NO_THP", "X3_TABLE_KEEP", "X3_PREPARE_SYNC"):
        os.environ.pop(k, None)
    os.environ.update(knobs)
    ts, fs = [], []
    for r in range(reps):
        t0 = time.perf_counter()
        L.x3_search_prepare(buf, n)
        while L.x3_search_ready() == 0:
            pass
        fs.append((time.perf_counter() - t0) * 1e3)
        L.x3_search_wait()
        ts.append((time.perf_counter() - t0) * 1e3)
        L.x3_search_release()
    print(knobs or "default", "ms per call:", " ".join(f"{t:.1f}" for t in ts), "| first piece:", " ".join(f"{t:.1f}" for t in fs), flush=True)
