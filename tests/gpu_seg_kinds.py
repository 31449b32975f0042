"""Segment search (or any variant) per data shape: 10 MB of each member kind of the C5 mix.
    python tests/gpu_seg_kinds.py [variant] [MB]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 6
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for kind in ("text", "exe", "img", "chem", "rec", "web", "xml", "mix"):
    data = np.frombuffer(corpus._member(kind, mb << 20, 500), dtype=np.uint8)
    best = 1e9
    for rep in range(3):
        ls, _, tm = pkg.search_host(data, W=8192, t=15, variant=variant, pinned=True)
        best = min(best, tm.kernel_ms)
    hist = np.bincount(ls, minlength=33)
    print(f"{kind:5s} {mb} MB: {best:8.3f} ms  {len(data) / best / 1e3:9.1f} MB/s   mean Lstar {ls.mean():.2f}  "
          f"share at 32: {hist[32] / len(ls):.3f}  at >= 8: {hist[8:].sum() / len(ls):.3f}", flush=True)
