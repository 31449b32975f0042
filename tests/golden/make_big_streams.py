"""Whole-stream KATs at BASELINE.json's sizes, from the COMPILED, UNMODIFIED reference
(oracle/_ref/x3_ref, built by oracle/Makefile from /root/reference).  Build container only:

    python tests/golden/make_big_streams.py [case ...]

Writes tests/golden/streams_big.json: length + sha256 of the reference `x3 -z` stream (and of
the input, so that a drifting generator is noticed) for

    C1 1 000 000 B              default flags  (BASELINE.json configs[0]; also decoded with x3 -d)
    C2 10 192 446 B             default flags  (configs[1], "byte-identical stream vs reference")
    C4 8 474 240 B              default flags  (configs[3])
    C5 scaled to 16 000 000 B   default flags  (configs[4] at a size the reference finishes here)
    C3 first 1 000 000 B        -w 1024 -t 64  (configs[2])

The reference needs ~10 s per MB at the default flags (single-threaded by design, x3.c:372-434), so
the cases run as parallel processes; the reference's own "elapsed time" is recorded beside each hash.
"""
import hashlib
import json
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

corpus = g.load_submodule("corpus")
X3 = ROOT / "oracle" / "_ref" / "x3_ref"
OUT = Path(__file__).parent / "streams_big.json"

CASES = [("C1", 1_000_000, "", True), ("C2", 10_192_446, "", False), ("C4", 8_474_240, "", False),
         ("C5", 16_000_000, "", False), ("C3", 1_000_000, "-w 1024 -t 64", False)]


def one(case):
    name, size, flags, decode = case
    data = corpus.generate(name, size)
    with tempfile.TemporaryDirectory() as td:
        src = Path(td) / "in.bin"
        dst = Path(td) / "out.x3"
        src.write_bytes(data)
        t0 = time.time()
        r = subprocess.run([str(X3), "-zf", *flags.split(), str(src), str(dst)], check=True,
                           stderr=subprocess.PIPE, text=True)
        wall = time.time() - t0
        el = [float(ln.split(":")[1]) for ln in r.stderr.splitlines() if ln.startswith("elapsed time:")]
        s = dst.read_bytes()
        rec = dict(len=len(s), sha256=hashlib.sha256(s).hexdigest(), in_sha256=hashlib.sha256(data).hexdigest(),
                   ref_elapsed_s=el[0] if el else None, ref_wall_s=round(wall, 1))
        if decode:
            back = Path(td) / "back.bin"
            subprocess.run([str(X3), "-df", str(dst), str(back)], check=True, stderr=subprocess.DEVNULL)
            rec["ref_round_trip"] = back.read_bytes() == data
    key = f"{name}:{size}:{flags}"
    print(key, rec, flush=True)
    return key, rec


def main():
    want = set(sys.argv[1:])
    cases = [c for c in CASES if not want or c[0] in want]
    streams = json.loads(OUT.read_text()) if OUT.exists() else {}
    with ThreadPoolExecutor(max_workers=len(cases)) as ex:
        for key, rec in ex.map(one, cases):
            streams[key] = rec
            OUT.write_text(json.dumps(streams, indent=1, sort_keys=True) + "\n")
    print("wrote", len(streams), "stream KATs to", OUT)


if __name__ == "__main__":
    main()
