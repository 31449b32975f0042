"""GPU parity tests of the segment search (x3_search_seg.cu, X3S_KERNEL_SEG -- what
X3S_KERNEL_DEFAULT runs at the reference's default flags): the table it returns through the C ABI
against the oracle (reference backend.c:58-78 restated in oracle/x3_oracle.c) on the same seeded
inputs, across input shapes, windows, thresholds, segment seams and the pipelined host path."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _inputs(corpus, kind, n):
    rng = np.random.Generator(np.random.PCG64(n + len(kind)))
    if kind in ("C1", "C2", "C4", "C5"):
        return np.frombuffer(corpus.generate(kind, n), dtype=np.uint8)
    if kind in ("text", "exe", "img", "chem", "rec", "web", "xml", "mix"):
        return np.frombuffer(corpus._member(kind, n, 500 + n % 7), dtype=np.uint8)
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


def _assert_seg_same(pkg, data, W, t, pinned=False, variant=None):
    variant = pkg.KERNEL_SEG if variant is None else variant
    lstar, _, tm = pkg.search_host(data, W=W, t=t, ngpus=1, variant=variant, pinned=pinned)
    _, ls_ref = ol.table(data, W, t)
    bad = np.nonzero(lstar != ls_ref)[0]
    assert len(bad) == 0, (f"segment search: Lstar differs at {len(bad)} positions, first {bad[:5].tolist()} "
                           f"got {lstar[bad[:5]].tolist()} oracle {ls_ref[bad[:5]].tolist()} (n={len(data)} W={W} t={t})")
    return tm


@pytest.mark.parametrize("kind,n", [("C1", 1), ("C1", 31), ("C1", 33), ("C1", 5000), ("C1", 30000), ("C4", 30000),
                                    ("C5", 60000), ("zeros", 40000), ("period", 30000), ("rand2", 30000),
                                    ("rand256", 30000), ("runs", 90000), ("C1", 200000), ("C5", 300000)])
def test_seg_search_equals_oracle_default_flags(pkg, corpus, kind, n):
    _assert_seg_same(pkg, _inputs(corpus, kind, n), 8192, 15)


@pytest.mark.parametrize("kind", ["text", "exe", "img", "chem", "rec", "web", "xml", "mix"])
def test_seg_search_equals_oracle_every_member_kind(pkg, corpus, kind):
    """every member shape of the C5 mix: the deep levels (chains, waves, whole-CTA groups) differ a lot between them"""
    _assert_seg_same(pkg, _inputs(corpus, kind, 150_000), 8192, 15)


@pytest.mark.parametrize("W", [34, 35, 64, 65, 100, 1024, 4096, 8191, 8193, 10000, 16384])
def test_seg_search_equals_oracle_window_sweep(pkg, corpus, W):
    _assert_seg_same(pkg, _inputs(corpus, "C1", 40000), W, 15)
    _assert_seg_same(pkg, _inputs(corpus, "rand2", 20000), W, 7)


@pytest.mark.parametrize("t", [5, 6, 15, 16, 30, 31, 32, 33, 64, 200, 254, 255, 300, 1000, 70000])
def test_seg_search_equals_oracle_threshold_sweep(pkg, corpus, t):
    _assert_seg_same(pkg, _inputs(corpus, "C5", 50000), 2048, t)
    _assert_seg_same(pkg, _inputs(corpus, "zeros", 30000), 1024, t)
    _assert_seg_same(pkg, _inputs(corpus, "C4", 40000), 8192, t)


def test_seg_search_segment_seams(pkg, corpus):
    """inputs a few positions around whole numbers of segments (B = 24 592 positions at W = 8192):
    the last segment is short, its window is the reference's zero padding (x3.c:579,590)"""
    B = 24592
    for n in (B - 1, B, B + 1, 2 * B, 2 * B + 17, 3 * B - 16):
        _assert_seg_same(pkg, _inputs(corpus, "C5", n), 8192, 15)
        _assert_seg_same(pkg, _inputs(corpus, "zeros", n), 8192, 15)


def test_seg_search_is_the_default_at_reference_flags(pkg, corpus):
    assert pkg.default_kernel(8192, 15, False) == pkg.KERNEL_SEG
    assert pkg.default_kernel(8192, 15, True) == pkg.KERNEL_STREAM        # the 32-bin table: brute force
    assert pkg.default_kernel(1 << 20, 64, False) == pkg.KERNEL_RANK      # C3's window does not fit on chip
    assert pkg.default_kernel(8192, 2, False) == pkg.KERNEL_RANK          # too many tiny groups for the queue
    data = _inputs(corpus, "C5", 400_000)
    tm = _assert_seg_same(pkg, data, 8192, 15, variant=pkg.KERNEL_DEFAULT)
    assert tm.launches == 1


@pytest.mark.parametrize("piece_mb", ["1", "2", None])
def test_seg_search_pipelined_pieces_do_not_change_results(pkg, corpus, monkeypatch, piece_mb):
    """between page-locked buffers a shard is uploaded, searched and copied back piece by piece
    (x3_search_api.cu); the pieces only change the queueing, never the table"""
    if piece_mb is not None:
        monkeypatch.setenv("X3_SEG_PIECE_MB", piece_mb)
    data = _inputs(corpus, "C5", 5_000_000)
    tm = _assert_seg_same(pkg, data, 8192, 15, pinned=True)
    if piece_mb is not None:
        assert tm.launches >= 2
    a, _, _ = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_SEG, pinned=False)
    b, _, _ = pkg.search_host(data, W=8192, t=15, variant=pkg.KERNEL_RANK, pinned=True)
    assert np.array_equal(a, b)


def test_seg_search_rejects_what_it_cannot_take(pkg):
    data = np.zeros(1000, dtype=np.uint8)
    for kw in (dict(W=8192, t=15, want_table=True), dict(W=1 << 20, t=15), dict(W=8192, t=2), dict(W=33, t=15)):
        with pytest.raises(pkg.X3SearchError) as ei:
            pkg.search_host(data, variant=pkg.KERNEL_SEG, **kw)
        assert ei.value.code == pkg.X3S_ERR_UNSUPP


@pytest.mark.parametrize("pinned", [False, True])
def test_streamed_table_lands_from_the_left(pkg, corpus, pinned, monkeypatch):
    """x3s_search_host_stream: while the call runs, *ready only grows, every position below it already holds
    its final Lstar, and it ends at n with the same table as x3s_search_host"""
    import ctypes as C
    import threading

    monkeypatch.setenv("X3_SEG_PIECE_MB", "1")
    data = _inputs(corpus, "C5", 6_000_000)
    n, W, t = len(data), 8192, 15
    want, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_SEG)
    L = pkg.lib()
    if pinned:
        px, pl = L.x3s_host_alloc(n + W), L.x3s_host_alloc(n)
        x = np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_uint8)), shape=(n + W,))
        out = np.ctypeslib.as_array(C.cast(pl, C.POINTER(C.c_uint8)), shape=(n,))
    else:
        x, out = np.zeros(n + W, dtype=np.uint8), np.empty(n, dtype=np.uint8)
    x[:n] = data
    x[n:] = 0
    out[:] = 255
    ready = C.c_size_t(0)
    rc = []
    th = threading.Thread(target=lambda: rc.append(L.x3s_search_host_stream(
        x.ctypes.data, n, W, t, 1, pkg.KERNEL_DEFAULT, out.ctypes.data, None, C.byref(ready))))
    th.start()
    seen, checks = 0, 0
    while th.is_alive():
        r = ready.value
        assert r >= seen
        if r > seen:
            assert np.array_equal(out[seen:r], want[seen:r])
            checks += 1
            seen = r
    th.join()
    assert rc == [0] and ready.value == n
    assert np.array_equal(out, want)
    assert checks >= 2          # it did arrive in pieces
    if pinned:
        L.x3s_host_free(px)
        L.x3s_host_free(pl)


def test_backend_prepare_returns_before_the_table_and_find_best_match_waits(pkg, corpus):
    """the backend.h drop-in: prepare starts the search and returns; find_best_match(p) answers once p has landed;
    x3_search_table() (which waits) shows the oracle's table"""
    import ctypes as C

    data = _inputs(corpus, "C5", 3_000_000)
    n, W, t = len(data), 8192, 15
    buf = pkg.padded(data, W)
    L = pkg.lib()
    L.set_forward_window(W)
    L.set_max_match_count(t)
    L.set_magic_factor1(4)
    L.set_magic_factor2(0)
    no_find = pkg.DICT_FIND_FN(lambda p: (1 << 64) - 1)
    no_len = pkg.DICT_LEN_FN(lambda i: 0)
    L.x3_backend_set_dict(C.cast(no_find, C.c_void_p), C.cast(no_len, C.c_void_p))
    try:
        L.x3_search_prepare(buf.ctypes.data, n)
        _, ls_ref = ol.table(data, W, t)
        L.find_best_match.restype = C.c_size_t
        L.find_best_match.argtypes = [C.c_void_p]
        for p in (0, 1, n // 2, n - 1):          # each waits for its piece; empty dictionary: the answer is max(Lstar, 1)
            assert L.find_best_match(buf.ctypes.data + p) == max(int(ls_ref[p]), 1)
        Lp, nn = C.c_void_p(), C.c_size_t()
        L.x3_search_table(None, C.byref(Lp), C.byref(nn))
        tab = np.ctypeslib.as_array(C.cast(Lp, C.POINTER(C.c_uint8)), shape=(nn.value,))
        assert nn.value == n and L.x3_search_ready() == n and np.array_equal(tab, ls_ref)
        assert L.x3_search_landed_ms() > 0
    finally:
        L.x3_search_release()
        L.x3_backend_set_dict(None, None)


@pytest.mark.parametrize("parts", [1, 2, 3, 8])
def test_pieces_dealt_out_in_turn_assemble_the_table(pkg, corpus, monkeypatch, parts):
    """x3s_search_host_part / x3s_search_device_part: ONE input, part p takes the pieces p, p + parts, ...; all parts
    together (here one after the other on one GPU, under torchrun one rank each) give the table of the plain search"""
    import ctypes as C
    import torch

    monkeypatch.setenv("X3_PART_PIECE_KB", "200")
    data = _inputs(corpus, "C5", 2_500_000)
    n, W, t = len(data), 8192, 15
    want, _, _ = pkg.search_host(data, W=W, t=t, variant=pkg.KERNEL_SEG)
    _, ls_ref = ol.table(data, W, t)
    assert np.array_equal(want, ls_ref)
    L = pkg.lib()
    piece = int(L.x3s_part_positions(W))
    assert piece > 0 and piece % 16 == 0 and (n + piece - 1) // piece >= 8
    x = pkg.padded(data, W)
    out = np.full(n, 255, dtype=np.uint8)
    tm = pkg.Timing()
    launches = 0
    for p in range(parts):
        assert L.x3s_search_host_part(x.ctypes.data, n, W, t, out.ctypes.data, C.byref(tm), p, parts) == 0, L.x3s_last_error()
        launches += tm.launches
        # only this part's pieces (and those of the parts before) are written so far
        done = np.zeros(n, dtype=bool)
        for q in range((n + piece - 1) // piece):
            if q % parts <= p:
                done[q * piece:(q + 1) * piece] = True
        assert np.array_equal(out[done], want[done]) and (out[~done] == 255).all()
    assert np.array_equal(out, want)
    # a part's pieces go in batches of 4, one launch each
    want_launches = 0
    for p in range(parts):
        m = len(range(p, (n + piece - 1) // piece, parts))
        want_launches += (m + 3) // 4
    assert launches == want_launches
    # device resident: the whole input in HBM, one launch per part
    dev = torch.device("cuda", 0)
    d_x = torch.zeros(pkg.required_bytes(n, W), dtype=torch.uint8, device=dev)
    d_x[:n].copy_(torch.from_numpy(data))
    d_l = torch.full((n,), 255, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream()
    for p in range(parts):
        assert L.x3s_search_device_part(0, d_x.data_ptr(), n, W, t, d_l.data_ptr(), s.cuda_stream, p, parts) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_l.cpu().numpy(), want)
    for bad in ((3, 3), (-1, 2), (0, 0)):
        assert L.x3s_search_host_part(x.ctypes.data, n, W, t, out.ctypes.data, None, bad[0], bad[1]) == pkg.X3S_ERR_ARG
    assert L.x3s_part_positions(1 << 20) == 0
    assert L.x3s_search_host_part(x.ctypes.data, n, 1 << 20, t, out.ctypes.data, None, 0, 1) == pkg.X3S_ERR_UNSUPP
