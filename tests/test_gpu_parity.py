"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI,
against the oracle on the same seeded inputs, against the golden fixtures, and -- at
BASELINE.json's full sizes -- through size-independent properties."""
import ctypes as C
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
GOLD = np.load(ROOT / "tests" / "golden" / "fbm_golden.npz")
KATS = json.loads((ROOT / "tests" / "golden" / "streams.json").read_text())


def _inputs(corpus, kind, n):
    rng = np.random.Generator(np.random.PCG64(n + len(kind)))
    if kind == "text":
        return np.frombuffer(corpus.generate("C1", n), dtype=np.uint8)
    if kind == "binary":
        return np.frombuffer(corpus.generate("C4", n), dtype=np.uint8)
    if kind == "mix":
        return np.frombuffer(corpus.generate("C5", n), dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "period":
        return (np.arange(n) % 7).astype(np.uint8)
    if kind == "rand2":
        return rng.integers(0, 2, n).astype(np.uint8)
    if kind == "rand256":
        return rng.integers(0, 256, n).astype(np.uint8)
    raise KeyError(kind)


def _assert_same(pkg, data, W, t, variant):
    lstar, H, _ = pkg.search_host(data, W=W, t=t, ngpus=1, variant=variant, want_table=True)
    H_ref, ls_ref = ol.table(data, W, t)
    if not np.array_equal(H, H_ref):
        bad = int(np.argmax((H != H_ref).any(axis=1)))
        raise AssertionError(f"H differs first at p={bad} (n={len(data)} W={W} t={t} variant={variant}):\n"
                             f" got {H[bad].tolist()}\n ref {H_ref[bad].tolist()}")
    assert np.array_equal(lstar, ls_ref), f"Lstar differs first at p={int(np.argmax(lstar != ls_ref))}"
    if variant in (0, 3):
        # production mode (no table): t <= 15 takes the saturating fast path of the stream kernel
        lstar2, _, _ = pkg.search_host(data, W=W, t=t, ngpus=1, variant=variant, want_table=False)
        assert np.array_equal(lstar2, ls_ref), (f"Lstar (no-table mode) differs first at "
                                                f"p={int(np.argmax(lstar2 != ls_ref))} (n={len(data)} W={W} t={t})")


@pytest.mark.parametrize("variant", [0, 2, 1])
@pytest.mark.parametrize("kind,n", [("text", 20000), ("binary", 9001), ("mix", 12345), ("zeros", 5000),
                                    ("period", 4097), ("rand2", 3968), ("rand256", 3969), ("text", 1),
                                    ("text", 31), ("text", 33)])
def test_table_equals_oracle_default_flags(pkg, corpus, variant, kind, n):
    _assert_same(pkg, _inputs(corpus, kind, n), 8192, 15, variant)


@pytest.mark.parametrize("variant", [0, 2, 1])
@pytest.mark.parametrize("W", [0, 1, 33, 34, 35, 64, 65, 100, 1024, 4096, 8191, 8193, 10000, 8192 + 8192, 17000,
                               65536])
def test_table_equals_oracle_window_sweep(pkg, corpus, variant, W):
    data = _inputs(corpus, "text", 6000)
    _assert_same(pkg, data, W, 15, variant)
    _assert_same(pkg, _inputs(corpus, "rand2", 4500), W, 3, variant)


@pytest.mark.parametrize("t", [-1 + 1, 1, 2, 3, 15, 16, 64, 200, 254])
def test_table_equals_oracle_threshold_sweep(pkg, corpus, t):
    for variant in (0, 2):
        _assert_same(pkg, _inputs(corpus, "mix", 9000), 2048, t, variant)
        _assert_same(pkg, _inputs(corpus, "zeros", 3000), 1024, t, variant)


def _assert_rank_same(pkg, data, W, t, pinned=False):
    """The rank search (variant 5, Lstar only) against the oracle."""
    lstar, _, _ = pkg.search_host(data, W=W, t=t, ngpus=1, variant=pkg.KERNEL_RANK, pinned=pinned)
    _, ls_ref = ol.table(data, W, t)
    assert np.array_equal(lstar, ls_ref), (f"rank search: Lstar differs first at p={int(np.argmax(lstar != ls_ref))} "
                                           f"(n={len(data)} W={W} t={t})")


@pytest.mark.parametrize("kind,n", [("text", 20000), ("binary", 9001), ("mix", 12345), ("zeros", 5000),
                                    ("period", 4097), ("rand2", 3968), ("rand256", 3969), ("text", 1),
                                    ("text", 31), ("text", 33), ("text", 70000), ("binary", 70000)])
def test_rank_search_equals_oracle_default_flags(pkg, corpus, kind, n):
    _assert_rank_same(pkg, _inputs(corpus, kind, n), 8192, 15)


@pytest.mark.parametrize("W", [0, 1, 33, 34, 35, 64, 65, 100, 1024, 4096, 8191, 8193, 10000, 17000, 65536, 1 << 20])
def test_rank_search_equals_oracle_window_sweep(pkg, corpus, W):
    n = 6000 if W <= 65536 else 1500
    _assert_rank_same(pkg, _inputs(corpus, "text", n), W, 15)
    _assert_rank_same(pkg, _inputs(corpus, "rand2", min(n, 4500)), W, 3)


@pytest.mark.parametrize("t", [0, 1, 2, 3, 15, 16, 64, 200, 254, 255, 256, 300, 1000, 70000])
def test_rank_search_equals_oracle_threshold_sweep(pkg, corpus, t):
    _assert_rank_same(pkg, _inputs(corpus, "mix", 9000), 2048, t)
    _assert_rank_same(pkg, _inputs(corpus, "zeros", 3000), 1024, t)
    _assert_rank_same(pkg, _inputs(corpus, "binary", 20000), 8192, t)


@pytest.mark.parametrize("kind,n,W,t", [("text", 30000, 8192, 15), ("zeros", 9000, 1024, 3), ("period", 20000, 8192, 15),
                                        ("binary", 40000, 8192, 15), ("rand2", 12000, 100, 50)])
def test_rank_search_without_tail_kernel(pkg, corpus, monkeypatch, kind, n, W, t):
    """X3_RANK_NO_TAIL keeps small arrays on the multi-launch path (level + radix kernels): both
    ways of finishing a search give the oracle's table."""
    data = _inputs(corpus, kind, n)
    monkeypatch.setenv("X3_RANK_NO_TAIL", "1")
    _assert_rank_same(pkg, data, W, t)
    monkeypatch.delenv("X3_RANK_NO_TAIL")
    _assert_rank_same(pkg, data, W, t)


@pytest.mark.parametrize("knob,value", [("X3_RANK_NO_PDL", "1"), ("X3_RANK_LAG", "1"), ("X3_RANK_LAG", "4"),
                                        ("X3_RANK_PROFILE", "1")])
def test_rank_search_launch_knobs_do_not_change_results(pkg, corpus, monkeypatch, knob, value):
    """Plain launches instead of programmatic dependent launch, other host lags, event-bracketed
    launches: only the queueing changes, never the table."""
    monkeypatch.setenv(knob, value)
    _assert_rank_same(pkg, _inputs(corpus, "text", 300_000), 8192, 15)
    _assert_rank_same(pkg, _inputs(corpus, "binary", 120_000), 4096, 40)
    _assert_rank_same(pkg, _inputs(corpus, "zeros", 30_000), 1024, 3)


@pytest.mark.parametrize("lanes,pieces", [(2, None), (3, None), (4, None), (2, "1"), (4, "1")])
def test_rank_search_lanes_do_not_change_results(pkg, corpus, monkeypatch, lanes, pieces):
    """The position range cut into lanes that are searched concurrently (own stream, own scratch, one
    host thread feeding all of them).  pieces=None: the host call is pipelined piece by piece (upload,
    search, copy back per piece); pieces='1': one upload, the device-level search forks into lanes and
    joins back.  Seams between lanes are invisible, also when a lane's window reaches over more than
    one neighbour (W = 64 KB against 75 000-position lanes)."""
    monkeypatch.setenv("X3_RANK_LANES", str(lanes))
    if pieces is not None:
        monkeypatch.setenv("X3_HOST_PIECES", pieces)
    pin = pieces is None  # the library pipelines a shard only between page-locked host buffers
    _assert_rank_same(pkg, _inputs(corpus, "text", 300_000), 8192, 15, pin)
    _assert_rank_same(pkg, _inputs(corpus, "binary", 280_000), 4096, 40, pin)
    _assert_rank_same(pkg, _inputs(corpus, "zeros", 270_000), 1024, 3, pin)
    _assert_rank_same(pkg, _inputs(corpus, "text", 300_000), 65536, 20, pin)
    _assert_rank_same(pkg, _inputs(corpus, "mix", 270_000), 300, 0, pin)   # t = 0: memset per lane
    _assert_rank_same(pkg, _inputs(corpus, "mix", 270_000), 33, 5, pin)    # empty window


def test_rank_search_pageable_buffers_are_not_pipelined(pkg, corpus, monkeypatch):
    """Pageable host buffers (what the backend.h shim gets from the reference's main) take the plain
    upload - search - copy back sequence whatever the knobs say; the table is the same."""
    monkeypatch.setenv("X3_RANK_LANES", "3")
    data = _inputs(corpus, "text", 300_000)
    a, _, tma = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK, pinned=False)
    b, _, tmb = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK, pinned=True)
    assert np.array_equal(a, b) and tma.launches == tmb.launches


def test_rank_search_lanes_default_one_per_chunk(pkg, corpus):
    """Without knobs an input of two chunks runs as two lanes (twice the set-up launches of one
    chunk, same table as the brute-force kernel)."""
    data = np.frombuffer(corpus.generate("C3", 17_400_000), dtype=np.uint8)
    ls_rank, _, tm = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK)
    ls_bf, _, _ = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_STREAM)
    assert np.array_equal(ls_rank, ls_bf), f"first difference at p={int(np.argmax(ls_rank != ls_bf))}"
    assert tm.launches >= 8


def test_rank_search_device_api_lanes(pkg, corpus, monkeypatch):
    """x3s_search_device on a caller's stream: the lanes fork from it and join back into it, so work
    queued on the stream behind the call sees the complete table."""
    import torch
    data = _inputs(corpus, "text", 400_000)
    W, t = 8192, 15
    _, ls_ref = ol.table(data, W, t)
    dev = torch.device("cuda:0")
    need = pkg.required_bytes(len(data), W)
    stream = torch.cuda.Stream(device=dev)
    for lanes in ("1", "3"):
        monkeypatch.setenv("X3_RANK_LANES", lanes)
        with torch.cuda.stream(stream):
            d_x = torch.zeros(need, dtype=torch.uint8, device=dev)
            d_x[: len(data)] = torch.from_numpy(np.array(data, copy=True)).to(dev, non_blocking=False)
            d_l = torch.full((len(data),), 99, dtype=torch.uint8, device=dev)
            pkg.search_device(0, d_x.data_ptr(), len(data), W, t, d_l.data_ptr(), None, stream.cuda_stream,
                              pkg.KERNEL_RANK)
            got = d_l.clone()  # queued on the same stream right behind the search
        stream.synchronize()
        assert np.array_equal(got.cpu().numpy(), ls_ref), f"lanes={lanes}"


def test_rank_search_rejects_table_request(pkg):
    with pytest.raises(pkg.X3SearchError):
        pkg.search_host(np.zeros(100, dtype=np.uint8), W=8192, t=15, variant=pkg.KERNEL_RANK, want_table=True)


@pytest.mark.parametrize("name,n,W,t", [("C2", 10_192_446, 8192, 15), ("C4", 8_474_240, 8192, 15),
                                        ("C3", 6_000_000, 65536, 64), ("C5", 20_000_000, 8192, 15)])
def test_rank_search_equals_brute_force_at_full_size(pkg, corpus, name, n, W, t):
    """Two independent algorithms (pair-test histogram vs occurrence rank) agree bit for bit
    on BASELINE.json-sized inputs; sampled bands also equal the oracle."""
    data = np.frombuffer(corpus.generate(name, n), dtype=np.uint8)
    ls_rank, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_RANK)
    ls_bf, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_STREAM)
    assert np.array_equal(ls_rank, ls_bf), f"first difference at p={int(np.argmax(ls_rank != ls_bf))}"
    for a in (0, len(data) - 4000):
        _, ls_ref = ol.table(data, W, t, p0=a, p1=a + 4000)
        assert np.array_equal(ls_rank[a:a + 4000], ls_ref)


def test_rank_search_chunked_input_equals_brute_force(pkg, corpus):
    """An input larger than one rank chunk (2^24 - D positions): chunk seams are invisible."""
    data = np.frombuffer(corpus.generate("C5", 36_000_000), dtype=np.uint8)
    ls_rank, _, _ = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK)
    ls_bf, _, _ = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_STREAM)
    assert np.array_equal(ls_rank, ls_bf), f"first difference at p={int(np.argmax(ls_rank != ls_bf))}"


def test_rank_search_chunk_seam_large_window(pkg, corpus):
    """A chunk seam with a 64 KB window and t = 20: the positions just in front of the seam take
    their followers from the next chunk's data (the trailing halo of the chunk)."""
    data = np.frombuffer(corpus.generate("C3", 17_400_000), dtype=np.uint8)
    W, t = 65536, 20
    ls_rank, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_RANK)
    ls_bf, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_STREAM)
    assert np.array_equal(ls_rank, ls_bf), f"first difference at p={int(np.argmax(ls_rank != ls_bf))}"


def test_rank_search_max_window_sample(pkg, corpus):
    """-w 1024 -t 64 (C3's flags): brute force needs 10^12 pair tests here, so sampled bands
    are checked against the oracle."""
    data = np.frombuffer(corpus.generate("C3", 3_000_000), dtype=np.uint8)
    W, t = 1 << 20, 64
    ls_rank, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_RANK)
    for a in (0, 1_500_000, len(data) - 3000):
        _, ls_ref = ol.table(data, W, t, p0=a, p1=a + 3000)
        assert np.array_equal(ls_rank[a:a + 3000], ls_ref)


def test_empty_and_rejected_inputs(pkg):
    ls, H, _ = pkg.search_host(np.zeros(0, dtype=np.uint8), W=8192, t=15, want_table=True)
    assert len(ls) == 0 and H.shape == (0, 32)
    with pytest.raises(pkg.X3SearchError) as ei:
        pkg.search_host(np.zeros(100, dtype=np.uint8), W=8192, t=255, want_table=True)  # u8 cells
    assert ei.value.code == pkg.X3S_ERR_UNSUPP


@pytest.mark.parametrize("key", [k for k in GOLD.files if k.startswith("fbm_") and k.endswith("_empty")])
def test_backend_mirror_matches_golden_empty_dict(pkg, key):
    """find_best_match() through the backend.h mirror == the reference's recorded values."""
    parts = key.split("_")
    name, W, t, f1, f2 = parts[1], int(parts[2][1:]), int(parts[3][1:]), int(parts[4][1:]), int(parts[5][1:])
    data = GOLD[f"in_{name}"]
    b = pkg.Backend()
    b.set_forward_window(W); b.set_max_match_count(t); b.set_magic_factor1(f1); b.set_magic_factor2(f2)
    no_find = ol.DICT_FIND_FN(lambda p: (1 << 64) - 1)
    no_len = ol.DICT_LEN_FN(lambda i: 0)
    b.set_dict(C.cast(no_find, C.c_void_p), C.cast(no_len, C.c_void_p))
    try:
        b.prepare(data)
        got = np.array([b.find_best_match(p) for p in range(len(data))], dtype=np.uint8)
    finally:
        b.release()
        b.set_dict(None, None)
        b.set_forward_window(8192); b.set_max_match_count(15); b.set_magic_factor1(4); b.set_magic_factor2(0)
    assert np.array_equal(got, GOLD[key])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not shipped")
def test_backend_mirror_matches_golden_live_dict(pkg):
    """Same with a populated dictionary: the filter (backend.c:79-90) runs against the
    compiled reference's dict.c, the histogram/selection come from the GPU."""
    R = ol.ref()
    data = GOLD["in_text3k"]
    x0 = ol.padded(data, 8192)
    if R.dict_get_elems() == 0:
        R.dict_enlarge()
        for off, ln in GOLD["dict_entries"]:
            assert ol.ref_dict_insert(x0, int(off), int(ln))
    b = pkg.Backend()
    b.set_dict(C.cast(R.dict_find_match, C.c_void_p), C.cast(R.dict_get_len_by_index, C.c_void_p))
    try:
        for key in [k for k in GOLD.files if k.endswith("_dict") and k.startswith("fbm_")]:
            parts = key.split("_")
            W, t, f1, f2 = int(parts[2][1:]), int(parts[3][1:]), int(parts[4][1:]), int(parts[5][1:])
            b.set_forward_window(W); b.set_max_match_count(t); b.set_magic_factor1(f1); b.set_magic_factor2(f2)
            b.prepare(data)
            got = np.array([b.find_best_match(p) for p in range(len(data))], dtype=np.uint8)
            assert np.array_equal(got, GOLD[key]), key
    finally:
        b.release()
        b.set_dict(None, None)
        b.set_forward_window(8192); b.set_max_match_count(15); b.set_magic_factor1(4); b.set_magic_factor2(0)


def test_full_size_c2_properties(pkg, corpus):
    """C2 at BASELINE.json's full size (10 192 446 B): the two independent kernels agree
    bit for bit, a sampled band equals the oracle, and halo sharding is invisible
    (searching a slice with its trailing halo reproduces the same rows)."""
    data = np.frombuffer(corpus.generate("C2"), dtype=np.uint8)
    assert len(data) == 10_192_446
    W, t = 8192, 15
    ls_bs, _, tm = pkg.search_host(data, W=W, t=t, variant=0)      # stream kernel, fast path
    ls_nv, _, _ = pkg.search_host(data, W=W, t=t, variant=1)       # naive byte loop
    assert np.array_equal(ls_bs, ls_nv)
    ls_full, _, _ = pkg.search_host(data, W=W, t=t, variant=4)     # stream kernel, u8 counters
    assert np.array_equal(ls_bs, ls_full)
    ls_v1, _, _ = pkg.search_host(data, W=W, t=t, variant=2)       # first bit-sliced kernel
    assert np.array_equal(ls_bs, ls_v1)
    for a in (0, 5_000_000, len(data) - 30000):
        _, ls_ref = ol.table(data, W, t, p0=a, p1=a + 30000)
        assert np.array_equal(ls_bs[a:a + 30000], ls_ref)
    # sharding property: rows [a, b) from the slice [a, b + W) with zero padding after the data
    a, bnd = 3_000_000 + 4096, 3_000_000 + 4096 + 250_000
    sl = data[a:bnd + W]
    ls_sl, _, _ = pkg.search_host(sl, W=W, t=t, variant=0)
    assert np.array_equal(ls_sl[: bnd - a], ls_bs[a:bnd])
    # checksum of the whole table pinned across variants
    assert hashlib.sha256(ls_bs.tobytes()).hexdigest() == hashlib.sha256(ls_nv.tobytes()).hexdigest()


def test_multi_gpu_sharding_equals_single(pkg, corpus):
    """x3s_search_host with ngpus > 1 returns the single-GPU table (needs a box with >= 2 GPUs: on one
    GPU "all visible" is one shard, and the comparison would prove nothing)."""
    if pkg.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    data = np.frombuffer(corpus.generate("C5", 3_000_000), dtype=np.uint8)
    one, _, tm1 = pkg.search_host(data, W=8192, t=15, ngpus=1)
    assert tm1.gpus == 1
    _, ls_ref = ol.table(data, 8192, 15, p0=1_400_000, p1=1_600_000)  # the band around the 2-GPU seam
    assert np.array_equal(one[1_400_000:1_600_000], ls_ref)
    for ngpus in sorted({2, min(3, pkg.device_count()), pkg.device_count()}):
        many, _, tm = pkg.search_host(data, W=8192, t=15, ngpus=ngpus)
        assert tm.gpus == ngpus >= 2
        assert np.array_equal(one, many), f"ngpus={ngpus}: first difference at p={int(np.argmax(one != many))}"
        pinned, _, tm = pkg.search_host(data, W=8192, t=15, ngpus=ngpus, pinned=True)
        assert tm.gpus == ngpus and np.array_equal(one, pinned)


@pytest.mark.parametrize("n,W,t", [(3 * 4096 + 17, 8192, 15), (2 * 4096, 65536, 3), (5 * 4096 + 1, 4096, 7),
                                   (70_000, 1 << 17, 20)])
def test_multi_gpu_seams_small_input_large_window(pkg, corpus, n, W, t):
    """Seam-focused: few positions, cuts at multiples of 4096, windows that reach over one or several
    neighbouring shards (the trailing halo is clamped at the end of the padded buffer)."""
    if pkg.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    data = _inputs(corpus, "text", n)
    _, ls_ref = ol.table(data, W, t)
    for ngpus in sorted({2, min(3, pkg.device_count()), pkg.device_count()}):
        got, _, tm = pkg.search_host(data, W=W, t=t, ngpus=ngpus)
        assert tm.gpus == min(ngpus, n // 4096 + 1)
        assert np.array_equal(got, ls_ref), f"ngpus={ngpus}: first difference at p={int(np.argmax(got != ls_ref))}"


@pytest.mark.skipif(not (REF / "x3_ref_dropin").exists(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("case", ["C1:60000:", "C4:30000:", "C5:40000:-n 3 -t 7", "C2:20000:-t 50 -w 32",
                                  "C1:60000:-m 1 -n 1", "C5:40000:-x"])
def test_reference_host_over_gpu_backend_matches_kat(case, corpus, tmp_path):
    """Drop-in proof: the UNMODIFIED reference host pass (x3.c, dict.c, ac.c ...) linked
    against libx3b200.so emits the reference's stream byte for byte."""
    name, size, flags = case.split(":")
    src = tmp_path / "in.bin"
    src.write_bytes(corpus.generate(name, int(size)))
    out = tmp_path / "out.x3"
    r = subprocess.run([str(REF / "x3_ref_dropin"), "-zf", *flags.split(), str(src), str(out)],
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    s = out.read_bytes()
    assert (len(s), hashlib.sha256(s).hexdigest()) == (KATS[case]["len"], KATS[case]["sha256"])
    subprocess.run([str(REF / "x3_ref"), "-df", str(out), str(tmp_path / "back")], check=True,
                   stderr=subprocess.DEVNULL)
    assert (tmp_path / "back").read_bytes() == src.read_bytes()


@pytest.mark.skipif(not (REF / "x3_ref_check").exists(), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("flags", ["", "-m 1 -n 1", "-t 3 -w 1"])
def test_every_call_equals_reference_on_gpu(flags, corpus, tmp_path):
    """Interposition harness over the GPU backend: every find_best_match() call of a real
    compress() compared with the compiled reference function, live dictionary."""
    src = tmp_path / "in.bin"
    src.write_bytes(corpus.generate("C1", 120000))
    r = subprocess.run([str(REF / "x3_ref_check"), "-zf", *flags.split(), str(src), str(tmp_path / "o.x3")],
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    assert "mismatches 0" in r.stderr
