"""Generates the golden fixtures of tests/golden/ from the COMPILED, UNMODIFIED
reference (oracle/_ref/libx3ref.so and oracle/_ref/x3_ref, built by oracle/Makefile
from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

fbm_golden.npz   reference find_best_match() (backend.c:56-100) for EVERY position of
                 small seeded inputs, under several (W, t, f1, f2), with an empty
                 dictionary and with a populated one.
streams.json     length + sha256 of the reference `x3 -z` stream for seeded inputs
                 and flag sets (whole-stream KATs).
"""
import hashlib
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g  # noqa: E402
import oracle_lib  # noqa: E402

corpus = g.load_submodule("corpus")

INPUTS = {
    "text3k": lambda: corpus.generate("C1", 3000),
    "bin3k": lambda: corpus.generate("C4", 3000),
    "mix3k": lambda: corpus.generate("C5", 3000)[512:3512],
    "runs2k": lambda: (b"abcabcabcabd" * 40 + b"\0" * 300 + b"xyzw" * 200 + bytes(range(256)) * 2)[:2000],
}
# (W, t, f1, f2)
PARAMS = [(8192, 15, 4, 0), (1024, 3, 4, 0), (100, 1, 4, 0), (34, 15, 4, 0), (33, 15, 4, 0), (0, 15, 4, 0),
          (4096, 50, 0, 0), (2048, 7, 1, 1), (8192, 15, 0, 3), (8192, 0, 4, 0), (65536, 64, 4, 0),
          # t >= 255: the reference takes any int (backend.c:21-26)
          (8192, 255, 4, 0), (8192, 300, 4, 0), (65536, 1000, 4, 0), (2048, 256, 0, 2)]


def fbm_all(R, data: bytes, W, t, f1, f2):
    x = oracle_lib.padded(data, W)
    R.set_forward_window(W)
    R.set_max_match_count(t)
    R.set_magic_factor1(f1)
    R.set_magic_factor2(f2)
    return np.array([R.find_best_match(x.ctypes.data + p) for p in range(len(data))], dtype=np.uint8), x


def main():
    R = oracle_lib.ref()
    out = {}
    meta = []
    # empty dictionary first (the dictionary of libx3ref.so is process-global and only grows)
    for name, fn in INPUTS.items():
        data = fn()
        out[f"in_{name}"] = np.frombuffer(data, dtype=np.uint8)
        for (W, t, f1, f2) in PARAMS:
            vals, _ = fbm_all(R, data, W, t, f1, f2)
            key = f"fbm_{name}_W{W}_t{t}_m{f1}_n{f2}_empty"
            out[key] = vals
            meta.append(key)
    # populated dictionary: strings taken from text3k at fixed offsets
    data = INPUTS["text3k"]()
    R.dict_enlarge()
    x0 = oracle_lib.padded(data, 8192)
    rng = np.random.Generator(np.random.PCG64(7))
    entries = []
    for _ in range(120):
        off = int(rng.integers(0, len(data) - 40))
        ln = int(rng.integers(1, 12))
        if oracle_lib.ref_dict_insert(x0, off, ln):
            entries.append((off, ln))
    out["dict_entries"] = np.array(entries, dtype=np.int64)
    for (W, t, f1, f2) in PARAMS:
        vals, _ = fbm_all(R, data, W, t, f1, f2)
        out[f"fbm_text3k_W{W}_t{t}_m{f1}_n{f2}_dict"] = vals
    np.savez_compressed(Path(__file__).parent / "fbm_golden.npz", **out)

    streams = {}
    x3 = ROOT / "oracle" / "_ref" / "x3_ref"
    cases = [("C1", 60000), ("C4", 30000), ("C5", 40000), ("C2", 20000)]
    flagsets = ["", "-t 1", "-t 3 -w 1", "-t 50 -w 32", "-m 0", "-m 1 -n 1", "-n 3 -t 7", "-x", "-w 0", "-t 0",
                "-t 255", "-t 300", "-t 1000 -w 64"]
    with tempfile.TemporaryDirectory() as td:
        for name, size in cases:
            data = corpus.generate(name, size)
            src = Path(td) / f"{name}.bin"
            src.write_bytes(data)
            for fl in flagsets:
                dst = Path(td) / "out.x3"
                subprocess.run([str(x3), "-zf", *fl.split(), str(src), str(dst)], check=True,
                               stderr=subprocess.DEVNULL)
                s = dst.read_bytes()
                streams[f"{name}:{size}:{fl}"] = dict(len=len(s), sha256=hashlib.sha256(s).hexdigest(),
                                                      in_sha256=hashlib.sha256(data).hexdigest())
    (Path(__file__).parent / "streams.json").write_text(json.dumps(streams, indent=1, sort_keys=True) + "\n")
    print("wrote", len(out), "arrays and", len(streams), "stream KATs")


if __name__ == "__main__":
    main()
