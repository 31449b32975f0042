"""GPU tests at BASELINE.json's sizes (run on the B200 box).

Tables: the production search over the whole of C2, C4 and C5 (default flags) against the sha256 of
the ORACLE's table (tests/golden/tables.json, computed by tests/golden/make_tables.py from
oracle/x3_oracle.c -- C5 is 212 M positions), plus oracle bands computed live; C3 at its real
flags (50 MB, -w 1024 -t 64: 5e13 byte compares, out of the oracle's reach as a whole) against the
recorded table and live oracle bands.
Streams: whole-stream KATs recorded from the compiled, unmodified reference
(tests/golden/streams_big.json, tests/golden/make_big_streams.py): C1 1 000 000 B, C2 whole,
C4 whole, C5 scaled to 16 MB, C3 first 1 MB at -w 1024 -t 64 -- through the product binary and,
for C1/C2, through the unmodified reference host pass linked against libx3b200.so."""
import hashlib
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
GPU_BIN = ROOT / "x3-compressor_b200" / "bin" / "x3"
TABLES = json.loads((ROOT / "tests" / "golden" / "tables.json").read_text())
BIG = json.loads((ROOT / "tests" / "golden" / "streams_big.json").read_text())


@pytest.mark.parametrize("name", ["C2", "C4", "C5", "C3"])
def test_production_table_at_full_size(pkg, corpus, name):
    if name not in TABLES:
        pytest.skip(f"no recorded table for {name}")
    rec = TABLES[name]
    data = np.frombuffer(corpus.generate_cached(name) if name == "C5" else corpus.generate(name), dtype=np.uint8)
    assert len(data) == rec["bytes"] and hashlib.sha256(data).hexdigest() == rec["in_sha256"], "generator drifted"
    W, t = rec["W"], rec["t"]
    ls, _, tm = pkg.search_host(data, W=W, t=t, ngpus=1, pinned=True)
    # live oracle bands: head, middle, tail (1 M positions in all at the default window)
    band = 350_000 if W <= 8192 else 2_000
    for a in (0, (len(data) // 2) & ~4095, len(data) - band):
        _, ls_ref = ol.table(data, W, t, p0=a, p1=a + band)
        assert np.array_equal(ls[a:a + band], ls_ref), f"{name}: band at {a} differs first at {a + int(np.argmax(ls[a:a + band] != ls_ref))}"
    assert hashlib.sha256(ls.tobytes()).hexdigest() == rec["lstar_sha256"], f"{name}: table differs from the {rec['source']} table"
    assert int(ls.sum(dtype=np.int64)) == rec["lstar_sum"]


def test_c2_whole_table_equals_oracle_live(pkg, corpus):
    """Every one of C2's 10 192 446 positions against the oracle computed in this test (no recorded
    hash in between): about half a minute of the box's host cores."""
    data = np.frombuffer(corpus.generate("C2"), dtype=np.uint8)
    ls, _, _ = pkg.search_host(data, W=8192, t=15, ngpus=1)
    _, ls_ref = ol.table(data, 8192, 15)
    assert np.array_equal(ls, ls_ref), f"first difference at p={int(np.argmax(ls != ls_ref))}"


def _kat(key):
    if key not in BIG:
        pytest.skip(f"{key} not recorded")
    return BIG[key]


def _run(binary, flags, src, out):
    r = subprocess.run([str(binary), "-zf", *flags.split(), str(src), str(out)], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    return out.read_bytes()


@pytest.mark.parametrize("key", ["C1:1000000:", "C2:10192446:", "C4:8474240:", "C5:16000000:", "C3:1000000:-w 1024 -t 64"])
def test_product_binary_whole_stream_kats(corpus, key, tmp_path):
    """bin/x3 -z emits the reference's stream (sha256 + length) at BASELINE.json's sizes, and its own
    decoder restores the input."""
    kat = _kat(key)
    name, size, flags = key.split(":")
    data = corpus.generate(name, int(size))
    assert hashlib.sha256(data).hexdigest() == kat["in_sha256"], "generator drifted"
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    s = _run(GPU_BIN, flags, src, tmp_path / "out.x3")
    assert (len(s), hashlib.sha256(s).hexdigest()) == (kat["len"], kat["sha256"]), key
    r = subprocess.run([str(GPU_BIN), "-df", str(tmp_path / "out.x3"), str(tmp_path / "back")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and (tmp_path / "back").read_bytes() == data


@pytest.mark.skipif(not (REF / "x3_ref_dropin").exists(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("key", ["C1:1000000:", "C2:10192446:"])
def test_reference_host_over_gpu_backend_whole_stream_kats(corpus, key, tmp_path):
    """configs[0] and configs[1]: the UNMODIFIED reference host pass over the GPU backend emits the
    reference's stream at full size; the reference's own x3 -d restores C1 (configs[0]'s round trip)."""
    kat = _kat(key)
    name, size, flags = key.split(":")
    data = corpus.generate(name, int(size))
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    s = _run(REF / "x3_ref_dropin", flags, src, tmp_path / "out.x3")
    assert (len(s), hashlib.sha256(s).hexdigest()) == (kat["len"], kat["sha256"]), key
    if name == "C1":
        subprocess.run([str(REF / "x3_ref"), "-df", str(tmp_path / "out.x3"), str(tmp_path / "back")], check=True,
                       stderr=subprocess.DEVNULL)
        assert (tmp_path / "back").read_bytes() == data


@pytest.mark.parametrize("flags", ["-t 255", "-t 300", "-t 1000 -w 64"])
def test_large_t_streams(corpus, flags, tmp_path):
    """-t >= 255 (the reference takes any int, backend.c:21-26): product binary and the reference host
    pass over the GPU backend emit the reference's streams (KATs in tests/golden/streams.json)."""
    kats = json.loads((ROOT / "tests" / "golden" / "streams.json").read_text())
    for name, size in (("C1", 60000), ("C4", 30000)):
        kat = kats[f"{name}:{size}:{flags}"]
        src = tmp_path / "in.bin"
        src.write_bytes(corpus.generate(name, size))
        for binary in (GPU_BIN, REF / "x3_ref_dropin"):
            if not binary.exists():
                continue
            s = _run(binary, flags, src, tmp_path / "out.x3")
            assert (len(s), hashlib.sha256(s).hexdigest()) == (kat["len"], kat["sha256"]), (name, flags, binary.name)
