"""The segment search's rules (tests/seg_model.py, the numpy statement of x3_search_seg.cu) against
the oracle: LSD levels 1-4 on shifted positions, rare first bytes, groups below t+2 dropped, the
kept rule, segment seams, virtual bytes in front of the input, the zero padding behind it."""
import numpy as np
import pytest

import oracle_lib as ol
import seg_model


def _data(corpus, kind, n):
    rng = np.random.Generator(np.random.PCG64(n))
    if kind == "text":
        return np.frombuffer(corpus.generate("C1", n), dtype=np.uint8)
    if kind == "mix":
        return np.frombuffer(corpus.generate("C5", n), dtype=np.uint8)
    if kind == "binary":
        return np.frombuffer(corpus.generate("C4", n), dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    if kind == "period":
        return (np.arange(n) % 7).astype(np.uint8)
    if kind == "rand2":
        return rng.integers(0, 2, n).astype(np.uint8)
    if kind == "runs":
        return np.frombuffer((b"abcabcabcabd" * 40 + b"\0" * 300 + b"xyzw" * 200 + bytes(range(256)) * 2) * 3, dtype=np.uint8)[:n]
    raise KeyError(kind)


@pytest.mark.parametrize("kind,n,W,t,m_max", [
    ("text", 20000, 8192, 15, 32768), ("text", 9000, 1024, 3, 4096), ("mix", 12000, 600, 7, 2048),
    ("binary", 8000, 2048, 15, 8192), ("zeros", 5000, 300, 15, 1024), ("period", 4097, 500, 5, 2048),
    ("rand2", 6000, 200, 50, 1024), ("runs", 5000, 1000, 2, 4096), ("text", 1, 8192, 15, 32768),
    ("text", 31, 8192, 15, 32768), ("text", 5000, 34, 1, 1024), ("text", 5000, 35, 15, 1024),
    ("mix", 7000, 8192, 1, 32768), ("zeros", 3000, 8192, 300, 32768), ("text", 6000, 4096, 255, 8192),
])
def test_segment_model_equals_oracle(corpus, kind, n, W, t, m_max):
    data = _data(corpus, kind, n)
    got = seg_model.lstar_segments(data, W, t, m_max=m_max)
    _, ref = ol.table(data, W, t)
    assert np.array_equal(got, ref), f"first difference at p={int(np.argmax(got != ref))}: {got[got != ref][:5]} vs {ref[got != ref][:5]}"


def test_segment_model_trivial_cases(corpus):
    data = _data(corpus, "text", 3000)
    for W, t in ((33, 15), (0, 15), (8192, 0), (8192, -1)):
        _, ref = ol.table(data, W, max(t, 0))
        assert np.array_equal(seg_model.lstar_segments(data, W, t), ref)
