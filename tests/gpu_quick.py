"""Small on-GPU sanity run (used under compute-sanitizer and for quick timing)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g  # noqa: E402
import oracle_lib as ol  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 1]
check = (len(sys.argv) <= 4) or sys.argv[4] != "nocheck"
data = np.frombuffer(corpus.generate("C2" if n > 1_000_000 else "C1", n), dtype=np.uint8)
for v in variants:
    for rep in range(3):
        ls, H, tm = pkg.search_host(data, W=W, t=15, variant=v, want_table=(check and rep == 0))
        print(f"variant {v} n={n} W={W} rep {rep}: h2d {tm.h2d_ms:.3f} kernel {tm.kernel_ms:.3f} d2h {tm.d2h_ms:.3f} "
              f"total {tm.total_ms:.3f} ms -> {n / tm.kernel_ms / 1e3:.2f} MB/s kernel", flush=True)
        if check and rep == 0:
            t0 = time.time()
            H_ref, ls_ref = ol.table(data, W, 15)
            ok = np.array_equal(H, H_ref) and np.array_equal(ls, ls_ref)
            print(f"  oracle check: {'OK' if ok else 'MISMATCH'} ({time.time() - t0:.1f}s oracle)")
            if not ok:
                bad = int(np.argmax((H != H_ref).any(axis=1)))
                print("  first bad position", bad, "\n  got", H[bad].tolist(), "\n  ref", H_ref[bad].tolist())
