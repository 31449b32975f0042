"""The product host program (x3-compressor_b200/host): CLI + sequential pass.

CPU part (`-m "not gpu"`): the host sources linked with the backend shim over the
oracle-backed fake device layer (oracle/build/x3_host_cpuoracle, test infrastructure)
must emit the reference's streams byte for byte (KATs recorded from the compiled,
unmodified reference in tests/golden/streams.json) and decode them again.
GPU part: the shipped binary bin/x3 over the real CUDA search against the same KATs.
"""
import hashlib
import json
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
KATS = json.loads((ROOT / "tests" / "golden" / "streams.json").read_text())
CPU_BIN = ROOT / "oracle" / "build" / "x3_host_cpuoracle"
GPU_BIN = ROOT / "x3-compressor_b200" / "bin" / "x3"
REF_BIN = ROOT / "oracle" / "_ref" / "x3_ref"
SMALL = [k for k in KATS if int(k.split(":")[1]) <= 100_000]
BIG = [k for k in KATS if int(k.split(":")[1]) > 100_000]


@pytest.fixture(scope="module")
def cpu_bin():
    if not CPU_BIN.exists():
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "hostcheck"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return CPU_BIN


def _roundtrip(binary, corpus, case, tmp_path, decoder=None):
    name, size, flags = case.split(":")
    data = corpus.generate(name, int(size))
    assert hashlib.sha256(data).hexdigest() == KATS[case]["in_sha256"], "generator drifted"
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    out = tmp_path / "out.x3"
    r = subprocess.run([str(binary), "-zf", *flags.split(), str(src), str(out)], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    s = out.read_bytes()
    assert (len(s), hashlib.sha256(s).hexdigest()) == (KATS[case]["len"], KATS[case]["sha256"]), case
    back = tmp_path / "back.bin"
    r = subprocess.run([str(decoder or binary), "-df", str(out), str(back)], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert back.read_bytes() == data
    return r.stderr


@pytest.mark.parametrize("case", SMALL)
def test_host_pass_emits_reference_stream_cpu(cpu_bin, corpus, case, tmp_path):
    _roundtrip(cpu_bin, corpus, case, tmp_path)


@pytest.mark.parametrize("threads", ["1", "2", "4"])
@pytest.mark.parametrize("case", SMALL[::4])
def test_host_pass_pipelines_emit_reference_stream_cpu(cpu_bin, corpus, case, threads, tmp_path, monkeypatch):
    """X3_THREADS = 1 (one thread), 2 (parse | code) and 4 (parse | ctx0 | ctx1 | code): every
    pipeline shape emits the reference's stream (an explicit X3_THREADS also applies to small inputs)."""
    monkeypatch.setenv("X3_THREADS", threads)
    _roundtrip(cpu_bin, corpus, case, tmp_path)


def test_host_pass_pipelines_agree_on_a_large_input(cpu_bin, corpus, tmp_path):
    """1.2 MB (many ring wrap-arounds, dictionary growth, big contexts): the three pipeline shapes
    emit one and the same stream, and it decodes back."""
    data = corpus.generate("C5", 1_200_000)
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    streams = []
    for threads in ("1", "2", "4"):
        out = tmp_path / f"out{threads}.x3"
        r = subprocess.run([str(cpu_bin), "-zf", str(src), str(out)], stderr=subprocess.PIPE, text=True,
                           env={**os.environ, "X3_THREADS": threads})
        assert r.returncode == 0, r.stderr[-800:]
        streams.append(out.read_bytes())
    assert streams[0] == streams[1] == streams[2]
    back = tmp_path / "back.bin"
    subprocess.run([str(cpu_bin), "-df", str(tmp_path / "out4.x3"), str(back)], check=True, stderr=subprocess.DEVNULL)
    assert back.read_bytes() == data


@pytest.mark.skipif(not REF_BIN.exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["C1:60000:", "C5:40000:-n 3 -t 7", "C4:30000:-x"])
def test_reference_decodes_our_stream(cpu_bin, corpus, case, tmp_path):
    _roundtrip(cpu_bin, corpus, case, tmp_path, decoder=REF_BIN)


def test_cli_behaviour(cpu_bin, corpus, tmp_path):
    """File-name handling, overwrite refusal and help of reference x3.c:484-548, file.c:47-55."""
    src = tmp_path / "f.txt"
    src.write_bytes(corpus.generate("C1", 5000))
    # one argument: adds .x3
    r = subprocess.run([str(cpu_bin), "-z", str(src)], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and (tmp_path / "f.txt.x3").exists()
    assert "Compressing..." in r.stderr and "forward window: 8192" in r.stderr and "max match count: 15" in r.stderr
    # refuses to overwrite without -f: abort()
    r = subprocess.run([str(cpu_bin), "-z", str(src)], stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0 and "File already exists" in r.stderr
    # one argument decompress: strips the suffix (needs -f because f.txt exists)
    first = (tmp_path / "f.txt").read_bytes()
    (tmp_path / "f.txt").unlink()
    r = subprocess.run([str(cpu_bin), "-d", str(tmp_path / "f.txt.x3")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and (tmp_path / "f.txt").read_bytes() == first
    assert "Decompressing..." in r.stderr
    # help
    r = subprocess.run([str(cpu_bin), "-h"], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "-w NUM : window size" in r.stderr
    # bad option aborts (x3.c:514-515)
    r = subprocess.run([str(cpu_bin), "-Q"], stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0
    # stdin/stdout with seekable files
    with open(src, "rb") as fi, open(tmp_path / "pipe.x3", "wb") as fo:
        r = subprocess.run([str(cpu_bin), "-z"], stdin=fi, stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0 and (tmp_path / "pipe.x3").read_bytes() == (tmp_path / "f.txt.x3").read_bytes()


def test_empty_and_tiny_inputs(cpu_bin, tmp_path):
    for data in (b"", b"a", b"ab", b"\0" * 40, bytes(range(256))):
        src = tmp_path / "t.bin"
        src.write_bytes(data)
        r = subprocess.run([str(cpu_bin), "-zf", str(src), str(tmp_path / "t.x3")], stderr=subprocess.PIPE)
        assert r.returncode == 0
        assert len((tmp_path / "t.x3").read_bytes()) % 4 == 0  # whole 32-bit words (bio.c:20-28)
        r = subprocess.run([str(cpu_bin), "-df", str(tmp_path / "t.x3"), str(tmp_path / "t.out")],
                           stderr=subprocess.PIPE)
        assert r.returncode == 0 and (tmp_path / "t.out").read_bytes() == data
        if REF_BIN.exists():
            subprocess.run([str(REF_BIN), "-zf", str(src), str(tmp_path / "r.x3")], stderr=subprocess.DEVNULL, check=True)
            assert (tmp_path / "r.x3").read_bytes() == (tmp_path / "t.x3").read_bytes()


def test_report_matches_reference(cpu_bin, corpus, tmp_path):
    """The stderr report (x3.c:662-693) carries the same numbers as the reference's."""
    if not REF_BIN.exists():
        pytest.skip("oracle/_ref not built")
    src = tmp_path / "in.bin"
    src.write_bytes(corpus.generate("C1", 60000))
    a = subprocess.run([str(cpu_bin), "-zf", str(src), str(tmp_path / "a.x3")], stderr=subprocess.PIPE, text=True).stderr
    b = subprocess.run([str(REF_BIN), "-zf", str(src), str(tmp_path / "b.x3")], stderr=subprocess.PIPE, text=True).stderr
    keep = ("input stream size", "output stream size", "dictionary:", "codestream size", "compression ratio",
            "number of events", "event sizes", "context entries")
    pick = lambda t: [ln for ln in t.splitlines() if ln.startswith(keep) or "compression ratio" in ln]
    assert pick(a) == pick(b)


def test_ratio_beyond_64_round_trips(cpu_bin, tmp_path):
    """300 KB of zeros compress 765:1.  The reference decoder allocates 64 x the stream size
    (x3.c:621) and cannot restore this; the product decoder grows its output on demand.  The stream
    itself is still the reference encoder's, byte for byte."""
    data = bytes(300_000)
    src = tmp_path / "z.bin"
    src.write_bytes(data)
    out = tmp_path / "z.x3"
    r = subprocess.run([str(cpu_bin), "-zf", str(src), str(out)], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    s = out.read_bytes()
    assert len(data) / len(s) > 64
    if REF_BIN.exists():
        ref = tmp_path / "z.ref.x3"
        subprocess.run([str(REF_BIN), "-zf", str(src), str(ref)], check=True, stderr=subprocess.DEVNULL)
        assert ref.read_bytes() == s
    back = tmp_path / "z.back"
    r = subprocess.run([str(cpu_bin), "-df", str(out), str(back)], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert back.read_bytes() == data


@pytest.mark.gpu
@pytest.mark.parametrize("case", SMALL[::3] + BIG)
def test_product_binary_emits_reference_stream_gpu(corpus, case, tmp_path):
    assert GPU_BIN.exists(), "x3-compressor_b200/bin/x3 missing: run __graft_entry__.build()"
    _roundtrip(GPU_BIN, corpus, case, tmp_path)


@pytest.mark.gpu
def test_product_binary_multi_gpu_env_does_not_change_stream(pkg, corpus, tmp_path):
    if pkg.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (X3_GPUS=0 means all visible: one GPU would compare 1 with 1)")
    src = tmp_path / "in.bin"
    src.write_bytes(corpus.generate("C5", 300000))
    outs = []
    for gpus in ("1", "0"):
        out = tmp_path / f"o{gpus}.x3"
        env = dict(os.environ, X3_GPUS=gpus)
        r = subprocess.run([str(GPU_BIN), "-zf", str(src), str(out)], stderr=subprocess.PIPE, env=env)
        assert r.returncode == 0
        outs.append(out.read_bytes())
    assert outs[0] == outs[1]
