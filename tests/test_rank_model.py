"""CPU tests of the occurrence-rank formulation (tests/rank_model.py) against the oracle:
the pruning rules and the rare-byte path of x3_search_rank.cu, checked without a GPU."""
import numpy as np
import pytest

import oracle_lib as ol
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
