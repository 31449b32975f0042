"""CPU tests of the oracle: the restatement against the golden vectors produced by
the compiled reference (tests/golden/make_golden.py), against the compiled reference
itself where oracle/_ref/ is present, and its two table builders against each other."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

GOLD = np.load(ol.ROOT / "tests" / "golden" / "fbm_golden.npz")


def _params(key):
    parts = key.split("_")
    return parts[1], int(parts[2][1:]), int(parts[3][1:]), int(parts[4][1:]), int(parts[5][1:]), parts[6]


EMPTY_KEYS = [k for k in GOLD.files if k.startswith("fbm_") and k.endswith("_empty")]
DICT_KEYS = [k for k in GOLD.files if k.startswith("fbm_") and k.endswith("_dict")]


@C.CFUNCTYPE(C.c_size_t, C.c_void_p)
def _no_dict_find(_p):
    return (1 << 64) - 1


@C.CFUNCTYPE(C.c_size_t, C.c_size_t)
def _no_dict_len(_i):
    return 0


@pytest.mark.parametrize("key", EMPTY_KEYS)
def test_restatement_matches_golden_empty_dict(key):
    """x3o_find_best_match (literal restatement of backend.c:56-100) and the split
    form (table -> Lstar -> filter) both reproduce the reference's return values."""
    name, W, t, f1, f2, _ = _params(key)
    data = GOLD[f"in_{name}"]
    want = GOLD[key]
    x = ol.padded(data, W)
    O = ol.oracle()
    got = np.array([O.x3o_find_best_match(x.ctypes.data + p, W, t, f1, f2, _no_dict_find, _no_dict_len)
                    for p in range(len(data))], dtype=np.uint8)
    assert np.array_equal(got, want)
    # split form: with an empty dictionary the filter never fires, so result = max(1, Lstar)
    H, ls = ol.table(data, W, t)
    assert np.array_equal(np.maximum(ls, 1), want)
    Hp, lsp = ol.table(data, W, t, plain=True)
    assert np.array_equal(H, Hp) and np.array_equal(ls, lsp)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_oracle_vs_compiled_reference_random():
    """Direct comparison with the compiled reference function on fresh random inputs."""
    R = ol.ref()
    if R.dict_get_elems() != 0:
        pytest.skip("reference dictionary already populated in this process")
    rng = np.random.Generator(np.random.PCG64(5))
    for W, t in [(8192, 15), (500, 2), (40, 1), (2000, 30)]:
        data = rng.integers(0, 5, 1500).astype(np.uint8)
        x = ol.padded(data, W)
        R.set_forward_window(W)
        R.set_max_match_count(t)
        R.set_magic_factor1(4)
        R.set_magic_factor2(0)
        want = np.array([R.find_best_match(x.ctypes.data + p) for p in range(len(data))], dtype=np.uint8)
        _, ls = ol.table(data, W, t)
        assert np.array_equal(np.maximum(ls, 1), want)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("key", DICT_KEYS)
def test_split_matches_golden_live_dict(key):
    """With a populated dictionary: Lstar + filter (backend.c:79-90) == reference.
    The dictionary is rebuilt inside the compiled reference's dict.c from the fixture."""
    name, W, t, f1, f2, _ = _params(key)
    data = GOLD[f"in_{name}"]
    want = GOLD[key]
    R = ol.ref()
    x = ol.padded(data, 8192 if W < 8192 else W)
    if R.dict_get_elems() == 0:
        R.dict_enlarge()
        x0 = ol.padded(data, 8192)
        test_split_matches_golden_live_dict.keep = x0
        for off, ln in GOLD["dict_entries"]:
            assert ol.ref_dict_insert(x0, int(off), int(ln))
    find = C.cast(R.dict_find_match, C.c_void_p)
    length = C.cast(R.dict_get_len_by_index, C.c_void_p)
    O = ol.oracle()
    H, ls = ol.table(data, W, t)
    got_split = np.array([O.x3o_filter_from_lstar(int(ls[p]), x.ctypes.data + p, f1, f2, find, length)
                          for p in range(len(data))], dtype=np.uint8)
    got_full = np.array([O.x3o_find_best_match(x.ctypes.data + p, W, t, f1, f2, find, length)
                         for p in range(len(data))], dtype=np.uint8)
    assert np.array_equal(got_full, want)
    assert np.array_equal(got_split, want)


@pytest.mark.parametrize("W", [0, 1, 32, 33, 34, 35, 64, 65, 100, 257, 1024])
@pytest.mark.parametrize("kind", ["rand4", "zeros", "period3", "rand256"])
def test_table_fast_equals_plain(W, kind):
    rng = np.random.Generator(np.random.PCG64(W * 7 + len(kind)))
    n = 700
    data = {"rand4": rng.integers(0, 4, n), "zeros": np.zeros(n), "period3": np.arange(n) % 3,
            "rand256": rng.integers(0, 256, n)}[kind].astype(np.uint8)
    for t in (0, 1, 15, 254):
        Hf, lf = ol.table(data, W, t)
        Hp, lp = ol.table(data, W, t, plain=True)
        assert np.array_equal(Hf, Hp)
        assert np.array_equal(lf, lp)
        H16, _ = ol.table(data, W, t, h16=True)
        assert np.array_equal(np.minimum(H16, 255).astype(np.uint8), Hf)


def test_histogram_semantics_edges():
    """backend.c:60-74: distances 1..W-33, zero padding participates, counts monotone."""
    data = np.zeros(50, dtype=np.uint8)
    x = ol.padded(data, 200)
    c = ol.histogram(x, 0, 200)
    assert (c == 200 - 33).all()           # all-zero data matches the zero padding
    assert (ol.histogram(x, 0, 33) == 0).all() and (ol.histogram(x, 0, 0) == 0).all()
    assert (ol.histogram(x, 0, 34) == 1).all()
    rng = np.random.Generator(np.random.PCG64(3))
    data = rng.integers(0, 3, 400).astype(np.uint8)
    x = ol.padded(data, 300)
    for p in (0, 17, 399):
        c = ol.histogram(x, p, 300)
        assert (np.diff(c.astype(np.int64)) <= 0).all()


def test_lstar_collapse_equals_loop_nest():
    """x3o_lstar_from_count == x3o_select with no dictionary, on random monotone counts."""
    rng = np.random.Generator(np.random.PCG64(11))
    O = ol.oracle()
    buf = np.zeros(64, dtype=np.uint8)
    for _ in range(3000):
        c = np.sort(rng.integers(0, rng.integers(1, 40), 32))[::-1].astype(np.uint64)
        t = int(rng.integers(-2, 30))
        cnt = ol.SIZE32(*[int(v) for v in c])
        sel = O.x3o_select(cnt, buf.ctypes.data, t, 4, 0, _no_dict_find, _no_dict_len)
        assert max(1, ol.lstar_from_count(c, t)) == sel
