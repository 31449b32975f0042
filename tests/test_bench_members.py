"""bench.py's weak-scaling members: member 0 is config C2, member r > 0 the same lines in a seeded
random order (same size, same bytes as a multiset, other positions), deterministic per rank."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def test_members_are_line_permutations_of_c2():
    corpus = g.load_submodule("corpus")
    m0 = bench.member(corpus, 0)
    assert m0.tobytes() == corpus.generate("C2") and len(m0) == bench.MEMBER_BYTES
    m1, m5 = bench.member(corpus, 1), bench.member(corpus, 5)
    for m in (m1, m5):
        assert len(m) == bench.MEMBER_BYTES
        assert np.array_equal(np.bincount(m, minlength=256), np.bincount(m0, minlength=256))
        assert sorted(m.tobytes().split(b"\n")) == sorted(m0.tobytes().split(b"\n"))
    assert not np.array_equal(m1, m0) and not np.array_equal(m1, m5)
    assert np.array_equal(bench.member(corpus, 1), m1)  # seeded: the same on every call and every rank


def test_padded_member_carries_the_next_members_head():
    corpus = g.load_submodule("corpus")
    x = bench.padded_member(corpus, 1, 3)
    assert len(x) == bench.MEMBER_BYTES + bench.W_BYTES
    assert np.array_equal(x[bench.MEMBER_BYTES:], bench.member(corpus, 2)[: bench.W_BYTES])
    last = bench.padded_member(corpus, 2, 3)
    assert not last[bench.MEMBER_BYTES:].any()  # the reference's zero padding behind the last member
