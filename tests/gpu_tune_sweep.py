"""Sweeps the rank search's launch-geometry knobs (X3_RANK_RSGRID, X3_RANK_LVGRID, X3_RANK_SMALL) on a few
inputs: best-of-5 search time per setting, table compared with the default's."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
cases = [("C2", 10_192_446), ("C4", 8_474_240), ("C3", 16_000_000)]
settings = [{}, {"X3_RANK_LVGRID": "6"}, {"X3_RANK_LVGRID": "5"}, {"X3_RANK_LVGRID": "12"}, {"X3_RANK_RSGRID": "5"}, {"X3_RANK_RSGRID": "6"},
            {"X3_RANK_RSGRID": "12"}, {"X3_RANK_SMALL": "1500000"}, {"X3_RANK_SMALL": "6000000"}, {"X3_RANK_SMALL": "12000000"}]
knobs = ("X3_RANK_RSGRID", "X3_RANK_LVGRID", "X3_RANK_SMALL")
for name, n in cases:
    data = np.frombuffer(corpus.generate(name, n), dtype=np.uint8)
    ref = None
    for st in settings:
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update(st)
        best = 1e9
        for rep in range(5):
            ls, _, tm = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK, pinned=True)
            best = min(best, tm.kernel_ms)
        if ref is None:
            ref = ls
        print(f"{name} {n} {st or 'default'}: {best:.3f} ms same={np.array_equal(ls, ref)}", flush=True)
