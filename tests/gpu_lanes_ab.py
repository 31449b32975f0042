"""Lanes / host pieces A/B on the GPU: the same inputs searched with 1..4 lanes, tables compared,
best-of-5 timings printed (device breakdown and the wall time of the whole host call).

    python tests/gpu_lanes_ab.py [quick]
"""
import os
import sys
import time
from pathlib import Path

import ctypes as C

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cases = [("C2", 10_192_446, 8192, 15), ("C4", 8_474_240, 8192, 15), ("C3", 20_000_000, 8192, 15)]
if not quick:
    cases += [("C5", 60_000_000, 8192, 15), ("C3", 12_000_000, 1 << 20, 64)]
bad = 0
for name, n, W, t in cases:
    data = np.frombuffer(corpus.generate(name, n), dtype=np.uint8)
    ref = None
    # pinned host buffers, as a host program that cares about the copies would use
    L = pkg.lib()
    hx = L.x3s_host_alloc(n + W)
    hl = L.x3s_host_alloc(n)
    xv = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_uint8)), shape=(n + W,))
    xv[:n] = data
    xv[n:] = 0
    lv = np.ctypeslib.as_array(C.cast(hl, C.POINTER(C.c_uint8)), shape=(n,))
    for lanes, pieces in ((1, None), (2, None), (3, None), (4, None), (2, 1), (3, 1), (4, 1)):
        # pieces=None: the host call is pipelined in `lanes` pieces; pieces=1: one upload, the device-level
        # search forks into `lanes` lanes, one copy back
        os.environ["X3_RANK_LANES"] = str(lanes)
        if pieces is None:
            os.environ.pop("X3_HOST_PIECES", None)
        else:
            os.environ["X3_HOST_PIECES"] = str(pieces)
        best = None
        wall = 1e9
        for rep in range(6):
            tm = pkg.Timing()
            lv[:] = 77
            t0 = time.perf_counter()
            rc = L.x3s_search_host(hx, n, W, t, 1, pkg.KERNEL_RANK, hl, None, C.byref(tm))
            wall = min(wall, (time.perf_counter() - t0) * 1e3)
            assert rc == 0, L.x3s_last_error()
            if best is None or tm.total_ms < best.total_ms:
                best = pkg.Timing.from_buffer_copy(tm)
        ls = lv.copy()
        if ref is None:
            ref = ls
        same = np.array_equal(ls, ref)
        bad += 0 if same else 1
        print(f"{name} n={n} W={W} t={t} lanes={lanes} pieces={pieces or lanes}: h2d {best.h2d_ms:.3f} kernel {best.kernel_ms:.3f} d2h {best.d2h_ms:.3f} "
              f"call {best.total_ms:.3f} ms ({n / best.total_ms / 1e3:.0f} MB/s) launches {best.launches} "
              f"same_as_1_lane={same}", flush=True)
    L.x3s_host_free(hx)
    L.x3s_host_free(hl)
print("FAILED" if bad else "ALL SAME", flush=True)
sys.exit(1 if bad else 0)
