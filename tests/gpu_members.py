"""Per-member-kind timing of the search kernel (data dependence of the rare-event path)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
for kind in ["text", "exe", "img", "chem", "rec", "web", "xml", "mix"]:
    data = np.frombuffer(corpus._member(kind, n, 77), dtype=np.uint8)
    best = 1e9
    for rep in range(3):
        ls, _, tm = pkg.search_host(data, W=W, t=15, variant=0)
        best = min(best, tm.kernel_ms)
    hist = np.bincount(ls, minlength=33)
    print(f"{kind:5s} n={n} W={W}: kernel {best:.3f} ms -> {n / best / 1e3:.1f} MB/s, {n * (W - 33) / best / 1e9:.2f} T pairs/s; "
          f"Lstar mean {ls.mean():.2f}, share Lstar>=9 {(ls >= 9).mean():.3f}, ==32 {(ls == 32).mean():.3f}", flush=True)
