"""CPU tests of the occurrence-rank formulation (tests/rank_model.py) against the oracle:
the pruning rules and the rare-byte path of x3_search_rank.cu, checked without a GPU."""
import numpy as np
import pytest

import oracle_lib as ol
import rank_model as rm
from rank_model import lstar_rank


def _data(corpus, kind, n):
    rng = np.random.Generator(np.random.PCG64(n))
    if kind in ("C1", "C3", "C4", "C5"):
        return np.frombuffer(corpus.generate(kind, n), dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "rand4":
        return rng.integers(0, 4, n).astype(np.uint8)
    if kind == "rand256":
        return rng.integers(0, 256, n).astype(np.uint8)
    if kind == "periodic":
        return np.tile(np.frombuffer(b"abcabcabd", dtype=np.uint8), n // 9 + 1)[:n]
    if kind == "runs":
        return np.repeat(rng.integers(0, 3, n // 20 + 1).astype(np.uint8), rng.integers(1, 60, n // 20 + 1))[:n]
    raise KeyError(kind)


@pytest.mark.parametrize("kind,n", [("C1", 9000), ("C4", 9000), ("C3", 7000), ("zeros", 1500), ("rand4", 6000),
                                    ("rand256", 6000), ("periodic", 5000), ("runs", 6000), ("C1", 1), ("C1", 40)])
@pytest.mark.parametrize("W,t", [(8192, 15), (1024, 1), (1024, 3), (40, 2), (34, 1), (33, 5), (100, 50),
                                 (20000, 64), (2048, 254), (300, 0)])
def test_rank_model_equals_oracle(corpus, kind, n, W, t):
    a = _data(corpus, kind, n)
    _, ref = ol.table(a, W, t)
    x = ol.padded(a, W)
    assert np.array_equal(lstar_rank(x, len(a), W, t), ref)
    # the kernel's tile-local participant rule (a superset) gives the same table
    assert np.array_equal(lstar_rank(x, len(a), W, t, tile=64), ref)
    # ... and so does keeping every element for level 2 (the kernels sort level 2 from x)
    assert np.array_equal(lstar_rank(x, len(a), W, t, tile=64, direct_level2=True), ref)


@pytest.mark.parametrize("seed,m,groups", [(1, 1, 1), (2, 5000, 1), (3, 5000, 7), (4, 20000, 900), (5, 3000, 3000)])
def test_merge_pass_equals_stable_sort(seed, m, groups):
    """The one-ranked-pass sort of DESIGN.md section 11 (byte pass, count matrix, block moves) gives exactly
    the stable (rank, byte) order the three radix passes give today."""
    rng = np.random.default_rng(seed)
    rank = np.sort(rng.integers(0, groups, m))               # the level kernel writes ranks in order
    byte = rng.choice(np.array([0, 32, 101, 116, 255]), m, p=[.1, .3, .3, .2, .1]) if seed % 2 else rng.integers(0, 256, m)
    nkey = rank.astype(np.int64) * 256 + byte
    assert np.array_equal(rm.merge_pass_order(nkey), np.argsort(nkey, kind="stable"))
