"""Small segment searches for compute-sanitizer (racecheck / memcheck / synccheck): inputs that reach every part of
the kernel -- LSD passes with the rare-byte map, chains, waves, whole-CTA groups, the jump -- checked against the oracle.
    compute-sanitizer --tool racecheck python tests/gpu_seg_sanitize.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g  # noqa: E402
import oracle_lib as ol  # noqa: E402

pkg = g.load_package()
corpus = g.load_submodule("corpus")
rng = np.random.Generator(np.random.PCG64(3))
cases = [("C5", np.frombuffer(corpus.generate("C5", 60000), dtype=np.uint8)),
         ("chem", np.frombuffer(corpus._member("chem", 70000, 501), dtype=np.uint8)),
         ("exe", np.frombuffer(corpus._member("exe", 70000, 502), dtype=np.uint8)),
         ("zeros", np.zeros(40000, dtype=np.uint8)),
         ("rand2", rng.integers(0, 2, 40000).astype(np.uint8))]
bad = 0
for name, data in cases:
    ls, _, tm = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_SEG)
    _, ref = ol.table(data, 8192, 15)
    ok = bool(np.array_equal(ls, ref))
    bad += not ok
    print(f"{name}: n={len(data)} kernel {tm.kernel_ms:.1f} ms  oracle check: {'OK' if ok else 'MISMATCH'}", flush=True)
# pieces in turn on one device
import ctypes as C
data = cases[0][1]
x = pkg.padded(data, 8192)
out = np.zeros(len(data), dtype=np.uint8)
import os
os.environ["X3_PART_PIECE_KB"] = "24"
for p in range(2):
    assert pkg.lib().x3s_search_host_part(x.ctypes.data, len(data), 8192, 15, out.ctypes.data, None, p, 2) == 0
_, ref = ol.table(data, 8192, 15)
print("pieces in turn:", "OK" if np.array_equal(out, ref) else "MISMATCH", flush=True)
sys.exit(1 if bad else 0)
