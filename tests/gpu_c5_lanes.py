"""C5 at full size (211.9 MB) with 1..8 lanes: search and end-to-end time from pinned buffers."""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
W, t = (1 << 20, 64) if name == "C3" else (8192, 15)
data = np.frombuffer(corpus.generate(name, 50_000_000) if name == "C3" else corpus.generate(name), dtype=np.uint8)
n = len(data)
L = pkg.lib()
hx, hl = L.x3s_host_alloc(n + W), L.x3s_host_alloc(n)
xv = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_uint8)), shape=(n + W,))
xv[:n] = data
xv[n:] = 0
lv = np.ctypeslib.as_array(C.cast(hl, C.POINTER(C.c_uint8)), shape=(n,))
ref = None
for lanes in (1, 2, 4, 6, 8):
    os.environ["X3_RANK_LANES"] = str(lanes)
    best = None
    for rep in range(3):
        tm = pkg.Timing()
        rc = L.x3s_search_host(hx, n, W, t, 1, pkg.KERNEL_DEFAULT, hl, None, C.byref(tm))
        assert rc == 0, L.x3s_last_error()
        if best is None or tm.total_ms < best.total_ms:
            best = pkg.Timing.from_buffer_copy(tm)
    if ref is None:
        ref = lv.copy()
    print(f"{name} {n} B lanes={lanes}: upload-to-first-search {best.h2d_ms:.2f} search {best.kernel_ms:.2f} tail copy {best.d2h_ms:.2f} "
          f"call {best.total_ms:.2f} ms ({n / best.total_ms / 1e3:.0f} MB/s) launches {best.launches} same={np.array_equal(lv, ref)}",
          flush=True)
