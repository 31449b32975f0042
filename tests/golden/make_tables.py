"""Lstar tables of the BASELINE.json configs from the ORACLE (oracle/x3_oracle.c, the CPU
restatement of reference backend.c:58-78 pinned against the compiled reference), as sha256 of
the whole table.  Build container only (minutes of CPU per config; C5 is 212 M positions):

    python tests/golden/make_tables.py C2 C4 C5

Writes tests/golden/tables.json.  C3 (-w 1024: 5e13 byte compares) is out of the oracle's reach;
its entry is the GPU table recorded in round 1 (both GPU formulations agreed on prefixes and the
oracle on bands) and is marked "gpu_recorded".
"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g  # noqa: E402
import oracle_lib  # noqa: E402

corpus = g.load_submodule("corpus")
OUT = Path(__file__).parent / "tables.json"
SLAB = 4_000_000


def main():
    names = sys.argv[1:] or ["C2", "C4"]
    tabs = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in names:
        cfg = corpus.CONFIGS[name]
        W, t = cfg["flags"]["w_kb"] * 1024, cfg["flags"]["t"]
        data = np.frombuffer(corpus.generate_cached(name), dtype=np.uint8)
        n = len(data)
        x = oracle_lib.padded(data, W)
        h = hashlib.sha256()
        tot = 0
        t0 = time.time()
        for p0 in range(0, n, SLAB):
            p1 = min(n, p0 + SLAB)
            ls = np.zeros(p1 - p0, dtype=np.uint8)
            oracle_lib.oracle().x3o_table_fast(x.ctypes.data, len(x), p0, p1, W, t, None, None, ls.ctypes.data)
            h.update(ls.tobytes())
            tot += int(ls.sum())
            print(name, p1, "of", n, f"{time.time() - t0:.0f} s", flush=True)
        tabs[name] = dict(bytes=n, W=W, t=t, lstar_sha256=h.hexdigest(), lstar_sum=tot, source="oracle",
                          in_sha256=hashlib.sha256(data.tobytes()).hexdigest(), oracle_s=round(time.time() - t0))
        OUT.write_text(json.dumps(tabs, indent=1, sort_keys=True) + "\n")
    print(json.dumps(tabs, indent=1))


if __name__ == "__main__":
    main()
