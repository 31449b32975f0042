"""On-GPU check of the rank search (variant 5) against the oracle (small inputs) and against
the brute-force stream kernel (large inputs), with kernel timings.

    python tests/gpu_rank_check.py [quick|full]
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
mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
rng = np.random.default_rng(7)
bad = 0


def report(name, n, W, t, ls, ref, tm, what):
    global bad
    ok = np.array_equal(ls, ref)
    msg = "OK" if ok else "MISMATCH"
    print(f"{name:10s} n={n:9d} W={W:8d} t={t:3d} vs {what:6s}: {msg}  kernel {tm.kernel_ms:9.3f} ms "
          f"({n / max(tm.kernel_ms, 1e-6) / 1e3:9.1f} MB/s) launches {tm.launches}", flush=True)
    if not ok:
        bad += 1
        i = int(np.argmax(ls != ref))
        print(f"    first bad position {i}: got {ls[i]} want {ref[i]}; {int((ls != ref).sum())} bad of {n}", flush=True)


small = []
for name, n in [("C1", 60000), ("C4", 60000), ("C3", 60000)]:
    small.append((name, np.frombuffer(corpus.generate(name, n), dtype=np.uint8)))
small.append(("zeros", np.zeros(20000, np.uint8)))
small.append(("rand4", rng.integers(0, 4, 50000).astype(np.uint8)))
small.append(("rand256", rng.integers(0, 256, 50000).astype(np.uint8)))
small.append(("periodic", np.tile(np.frombuffer(b"abcabcabd", dtype=np.uint8), 5000)))
small.append(("runs", np.repeat(rng.integers(0, 3, 900).astype(np.uint8), rng.integers(1, 120, 900))))
small.append(("tiny", np.frombuffer(b"abracadabra abracadabra", dtype=np.uint8)))
small.append(("one", np.frombuffer(b"x", dtype=np.uint8)))
flags = [(8192, 15), (1024, 1), (1024, 3), (40, 2), (34, 1), (33, 5), (100, 50), (65536, 64), (8192, 254), (300, 0)]
if mode == "quick":
    flags = [(8192, 15), (1024, 1), (40, 2), (34, 1), (65536, 64), (8192, 254)]
for name, a in small:
    for (W, t) in flags:
        _, ref = ol.table(a, W, t)
        ls, _, tm = pkg.search_host(a, W=W, t=t, ngpus=1, variant=pkg.KERNEL_RANK)
        report(name, len(a), W, t, ls, ref, tm, "oracle")

big = [("C2", 10_192_446, 8192, 15), ("C4", 8_474_240, 8192, 15), ("C3", 4_000_000, 65536, 64),
       ("C3", 20_000_000, 8192, 15)]
if mode == "full":
    big += [("C3", 50_000_000, 1 << 20, 64), ("C5", 40_000_000, 8192, 15)]
for name, n, W, t in big:
    a = np.frombuffer(corpus.generate(name, n), dtype=np.uint8)
    best = None
    for rep in range(3):
        ls, _, tm = pkg.search_host(a, W=W, t=t, ngpus=1, variant=pkg.KERNEL_RANK)
        if best is None or tm.kernel_ms < best.kernel_ms:
            best = tm
    if W <= 65536:
        ref, _, tms = pkg.search_host(a, W=W, t=t, ngpus=1, variant=pkg.KERNEL_STREAM)
        report(name, n, W, t, ls, ref, best, "stream")
        print(f"    stream kernel: {tms.kernel_ms:.3f} ms", flush=True)
    else:
        # brute force would take seconds: check a sample of positions against the oracle
        m = 40_000
        for p0 in (n // 2, n - m):
            _, ref = ol.table(a, W, t, p0=p0, p1=p0 + m)
            report(name, n, W, t, ls[p0:p0 + m], ref, best, "oracle")
print("FAILED" if bad else "ALL OK", flush=True)
sys.exit(1 if bad else 0)
