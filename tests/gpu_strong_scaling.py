"""One input, G GPUs of one process (x3s_search_host halo sharding, pinned host buffers):
end-to-end time of the C ABI call against the number of GPUs.

    python tests/gpu_strong_scaling.py [config] [bytes]
"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
size = int(sys.argv[2]) if len(sys.argv) > 2 else None
W, t = 8192, 15
data = np.frombuffer(corpus.generate(cfg, size) if size else corpus.generate(cfg), dtype=np.uint8)
n = len(data)
L = pkg.lib()
hx = L.x3s_host_alloc(n + W)
hl = L.x3s_host_alloc(n)
C.memset(hx, 0, n + W)
C.memmove(hx, data.ctypes.data, n)
ref = None
for G in [x for x in (1, 2, 4, 8) if x <= pkg.device_count()]:
    tm = pkg.Timing()
    best = 1e9
    for rep in range(6):
        t0 = time.perf_counter()
        rc = L.x3s_search_host(hx, n, W, t, G, pkg.KERNEL_DEFAULT, hl, None, C.byref(tm))
        dt = time.perf_counter() - t0
        assert rc == 0, L.x3s_last_error()
        if rep >= 2:
            best = min(best, dt)
    ls = np.ctypeslib.as_array(C.cast(hl, C.POINTER(C.c_uint8)), shape=(n,)).copy()
    if ref is None:
        ref = ls
    same = bool(np.array_equal(ls, ref))
    print(f"{cfg} {n} B on {G} GPU(s): {best * 1e3:.3f} ms end to end -> {n / best / 1e6:.0f} MB/s "
          f"(kernel {tm.kernel_ms:.3f} ms max over GPUs), table identical to 1 GPU: {same}", flush=True)
