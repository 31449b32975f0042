"""Small on-GPU sanity run (used under compute-sanitizer, ncu and for quick timing).

    python tests/gpu_quick.py [n] [W] [variants] [check|nocheck] [config] [t]
"""
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
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2, 1]
check = (len(sys.argv) <= 4) or sys.argv[4] != "nocheck"
cfg = sys.argv[5] if len(sys.argv) > 5 else ("C2" if n > 1_000_000 else "C1")
t = int(sys.argv[6]) if len(sys.argv) > 6 else 15
if cfg == "zeros":
    data = np.zeros(n, dtype=np.uint8)
else:
    data = np.frombuffer(corpus.generate(cfg, n), dtype=np.uint8)
ref = None
for v in variants:
    for rep in range(3):
        want = check and rep == 0 and v not in (5, 6)  # the rank and segment searches return Lstar only
        ls, H, tm = pkg.search_host(data, W=W, t=t, variant=v, want_table=want)
        print(f"variant {v} {cfg} n={n} W={W} t={t} rep {rep} table={want}: h2d {tm.h2d_ms:.3f} kernel {tm.kernel_ms:.3f} "
              f"d2h {tm.d2h_ms:.3f} total {tm.total_ms:.3f} ms -> {n / tm.kernel_ms / 1e3:.2f} MB/s, "
              f"{n * max(W - 33, 0) / tm.kernel_ms / 1e9:.2f} T pairs/s", flush=True)
        if check:
            if ref is None:
                t0 = time.time()
                ref = ol.table(data, W, t)
                print(f"  (oracle {time.time() - t0:.1f}s)")
            H_ref, ls_ref = ref
            ok = np.array_equal(ls, ls_ref) and (H is None or np.array_equal(H, H_ref))
            print(f"  oracle check: {'OK' if ok else 'MISMATCH'}")
            if not ok:
                if H is not None and not np.array_equal(H, H_ref):
                    bad = int(np.argmax((H != H_ref).any(axis=1)))
                    print("  first bad H position", bad, "\n  got", H[bad].tolist(), "\n  ref", H_ref[bad].tolist())
                    print("  bad rows:", int((H != H_ref).any(axis=1).sum()))
                if not np.array_equal(ls, ls_ref):
                    bad = int(np.argmax(ls != ls_ref))
                    print("  first bad Lstar position", bad, "got", ls[bad], "ref", ls_ref[bad], "count bad", int((ls != ls_ref).sum()))
