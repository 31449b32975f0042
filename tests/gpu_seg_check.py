"""First-contact check of the segment search (variant 6) on the GPU box: a ladder of cases against
the oracle with the first mismatches printed, then timings at BASELINE sizes against the recorded
table hashes.

    python tests/gpu_seg_check.py [small|big|all]
"""
import hashlib
import json
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
what = sys.argv[1] if len(sys.argv) > 1 else "all"
bad_total = 0


def data_of(kind, n):
    rng = np.random.Generator(np.random.PCG64(n))
    if kind in ("C1", "C2", "C4", "C5"):
        return np.frombuffer(corpus.generate(kind, n), dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "period":
        return (np.arange(n) % 7).astype(np.uint8)
    if kind == "rand2":
        return rng.integers(0, 2, n).astype(np.uint8)
    if kind == "rand256":
        return rng.integers(0, 256, n).astype(np.uint8)
    if kind == "runs":
        return np.frombuffer((b"abcabcabcabd" * 40 + b"\0" * 3000 + b"xyzw" * 2000 + bytes(range(256)) * 20) * 40, dtype=np.uint8)[:n]
    raise KeyError(kind)


def check(kind, n, W, t, pinned=False):
    global bad_total
    data = data_of(kind, n)
    t0 = time.time()
    ls, _, tm = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_SEG, pinned=pinned)
    _, ref = ol.table(data, W, t)
    bad = np.nonzero(ls != ref)[0]
    print(f"{kind:8s} n={n:8d} W={W:6d} t={t:5d} pinned={int(pinned)}: kernel {tm.kernel_ms:8.3f} ms  "
          f"{'OK' if len(bad) == 0 else f'MISMATCH at {len(bad)} positions, first {bad[:6].tolist()} got {ls[bad[:6]].tolist()} ref {ref[bad[:6]].tolist()}'}"
          f"  ({time.time() - t0:.1f} s)", flush=True)
    bad_total += len(bad)


if what in ("small", "all"):
    for (kind, n) in (("C1", 1), ("C1", 31), ("C1", 5000), ("C1", 30000), ("C4", 30000), ("C5", 60000), ("zeros", 40000),
                      ("period", 30000), ("rand2", 30000), ("rand256", 30000), ("runs", 90000), ("C1", 200000), ("C5", 300000)):
        check(kind, n, 8192, 15)
    for W in (34, 35, 64, 100, 1024, 4096, 8191, 8193, 10000, 16384):
        check("C1", 40000, W, 15)
        check("rand2", 20000, W, 7)
    for t in (5, 6, 15, 16, 30, 31, 32, 33, 64, 200, 254, 255, 300, 1000, 70000):
        check("C5", 50000, 2048, t)
        check("zeros", 30000, 1024, t)
        check("C4", 40000, 8192, t)
    check("C5", 3_000_000, 8192, 15, pinned=True)
if what in ("big", "all"):
    tabs = json.loads((ROOT / "tests" / "golden" / "tables.json").read_text())
    for name in ("C2", "C4", "C5"):
        data = np.frombuffer(corpus.generate_cached(name) if name == "C5" else corpus.generate(name), dtype=np.uint8)
        for variant, label in ((pkg.KERNEL_SEG, "seg "), (pkg.KERNEL_RANK, "rank")):
            best = None
            for rep in range(4):
                ls, _, tm = pkg.search_host(data, W=8192, t=15, variant=variant, pinned=True)
                if best is None or tm.kernel_ms < best.kernel_ms:
                    best = pkg.Timing.from_buffer_copy(tm)
            sha = hashlib.sha256(ls.tobytes()).hexdigest()
            ok = sha == tabs[name]["lstar_sha256"]
            print(f"{name} {label}: kernel {best.kernel_ms:8.3f} ms  call {best.total_ms:8.3f} ms  {len(data) / best.kernel_ms / 1e3:9.1f} MB/s  "
                  f"launches {best.launches}  table {'== recorded' if ok else 'DIFFERS from recorded'}", flush=True)
            if not ok:
                bad_total += 1
                band = 400000
                _, ref = ol.table(data, 8192, 15, p0=0, p1=band)
                bad = np.nonzero(ls[:band] != ref)[0]
                print(f"   first band: {len(bad)} mismatches, first {bad[:8].tolist()} got {ls[bad[:8]].tolist()} ref {ref[bad[:8]].tolist()}")
print("TOTAL MISMATCHES", bad_total)
sys.exit(1 if bad_total else 0)
