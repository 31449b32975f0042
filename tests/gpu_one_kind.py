"""One member kind of the C5 mix through the search (for ncu captures and quick timings).
    python tests/gpu_one_kind.py <kind> [MB] [reps] [variant]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
kind = sys.argv[1]
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 6
data = np.frombuffer(corpus._member(kind, mb << 20, 500), dtype=np.uint8)
for rep in range(reps):
    ls, _, tm = pkg.search_host(data, W=8192, t=15, variant=variant, pinned=True)
    print(f"{kind} {mb} MB rep {rep}: kernel {tm.kernel_ms:.3f} ms, launches {tm.launches}", flush=True)
