"""numpy/Python statement of the window-resident ("segment") search of x3_search_seg.cu: what the
kernel computes, without its warp mechanics.  TEST INFRASTRUCTURE: pins the algorithm's rules
against the oracle on the CPU (tests/test_seg_model.py) before any GPU is involved.

One segment = B searched positions [a, a+B) plus everything they can see: the elements are the
positions p in [a-3, a+B+D) (local id e = p - (a-3)), M = B + D + 3 of them, all held on chip.

  levels 1..4  four stable counting sorts on the bytes x[p+3], x[p+2], x[p+1], x[p] (LSD).  After
               the pass on x[p+j] the elements are ordered by (x[p+j..p+3], p), which is the
               level-(4-j) order of the positions q = p + j, so each pass's output is tested at
               one level more: passed(i) <=> element i+t+1 has the same gram and lies within D.
               Positions that fail level 1 have c1 <= t followers: Lstar = min LCP32 over them
               (0 when c1 < 2) -- backend.c:76-78 collapsed for tc* = c1 - 1.
  levels 4..32 groups (runs of equal 4-grams, then split by the next byte) with fewer than t+2
               elements can never pass and are dropped; a group is tested, pruned to the elements
               within D behind a passed one (the only followers that can matter deeper down) and
               split by x[p+L] into the level-(L+1) groups.
  Lstar[q]     = the deepest level q passed.
Elements in front of the input (p < 0, first segment) carry virtual zero bytes: they sort to the
front of their groups and are never a follower of a real position, nor searched themselves.
"""
from __future__ import annotations

import numpy as np


def lstar_segments(data: np.ndarray, W: int, t: int, m_max: int = 32768, b_cap: int | None = None) -> np.ndarray:
    n = len(data)
    out = np.zeros(n, dtype=np.uint8)
    D = W - 33 if W > 33 else 0
    if t <= 0 or D == 0 or n == 0:
        return out  # backend.c:76 never enters the selection / empty window: find_best_match returns 1
    B = (m_max - D - 3) // 16 * 16
    if b_cap is not None:
        B = min(B, b_cap)
    assert B >= 16, "window too large for a resident segment"
    x = np.zeros(n + W + 64, dtype=np.uint8)
    x[:n] = data
    for a in range(0, n, B):
        _segment(x, n, a, min(B, n - a), D, t, out)
    return out


def _segment(x, n, a, Bs, D, t, out):
    la = t + 1
    M = min(Bs + D + 3, n + D + 3 - a)          # elements e in [0, M): positions p = a - 3 + e < n + D
    xs = np.zeros(M + 64, dtype=np.uint8)        # xs[k] = x[a - 3 + k], zero in front of the input
    lo = a - 3
    src0 = max(lo, 0)
    xs[src0 - lo: src0 - lo + (M + 48 - (src0 - lo))] = x[src0: src0 + (M + 48 - (src0 - lo))]
    L8 = np.zeros(Bs, dtype=np.uint8)
    order = np.arange(M, dtype=np.int64)
    for k in range(1, 5):                        # pass k sorts on byte 4-k; its output is tested at level k
        dig = xs[order + (4 - k)]
        order = order[np.argsort(dig, kind="stable")]
        gram = np.stack([xs[order + j] for j in range(4 - k, 4)], axis=1)      # the k bytes that define the groups
        same = np.zeros(M, dtype=bool)
        if M > la:
            same[: M - la] = (gram[: M - la] == gram[la:]).all(axis=1) & (order[la:] - order[: M - la] <= D)
        q = order + 1 - k                        # searched position (relative to a) this element stands for
        subj = (q >= 0) & (q < Bs)
        L8[q[subj & same]] = k
        if k == 1:
            # the rare first bytes: fewer than t+1 followers within D
            for i in np.nonzero(subj & ~same)[0]:
                c1, best = 0, 32
                j = i + 1
                while j < M and c1 < t and gram[j, 0] == gram[i, 0] and order[j] - order[i] <= D:
                    c1 += 1
                    pa, pb = order[i] + 3, order[j] + 3
                    l = 0
                    while l < 32 and xs[pa + l] == xs[pb + l]:
                        l += 1
                    best = min(best, l)
                    j += 1
                L8[q[i]] = best if c1 >= 2 else 0
        if k == 4:
            passed4 = subj & same
    # level-4 groups; the generic step from here on: (elements of one group in position order, level L)
    heads = np.ones(M, dtype=bool)
    heads[1:] = (gram[1:] != gram[:-1]).any(axis=1)
    starts = np.nonzero(heads)[0]
    ends = np.append(starts[1:], M)
    work = [(order[s:e], 4) for s, e in zip(starts, ends) if e - s >= t + 2]
    while work:
        el, L = work.pop()
        g = len(el)
        q = el - 3
        passed = np.zeros(g, dtype=bool)
        if g > la:
            passed[: g - la] = (el[la:] - el[: g - la] <= D)
        passed &= (q >= 0) & (q < Bs)
        if not passed.any():
            continue
        L8[q[passed]] = L
        if L == 32:
            continue
        # kept: within D behind the last passed element at or in front of it
        lastp = np.maximum.accumulate(np.where(passed, el, -(1 << 40)))
        kept = el - lastp <= D
        ke = el[kept]
        by = xs[ke + L]
        for b in np.unique(by):
            sub = ke[by == b]
            if len(sub) >= t + 2:
                work.append((sub, L + 1))
    out[a:a + Bs] = L8
