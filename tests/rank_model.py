"""numpy model of the occurrence-rank search (x3-compressor_b200/csrc/x3_search_rank.cu).

TEST INFRASTRUCTURE.  The model states the algorithm the rank kernels implement, level by
level, so that its pruning rules can be checked against the oracle on a machine without a
GPU (tests/test_rank_model.py).  It computes Lstar only:

    count_L(p) = #{ q in (p, p+D] : x[q..q+L) == x[p..p+L) },  D = W - 33   (backend.c:58-74)
    tc*(p)     = min(t, count_1(p) - 1)                                   (backend.c:76-78)
    Lstar(p)   = #{ L : count_L(p) > tc*(p) }   (0 when t <= 0 or count_1(p) < 2)

Level L keeps an array of positions sorted by (L-gram, position).  In that order
"count_L(p) > t" is one lookup: the element t+1 places further on has the same L-gram and
lies within D of p.  Positions whose first byte occurs at most t times in their window
(tc* < t) are settled at level 1 by walking their <= t followers.  Between levels the array
is pruned to the elements that can still matter: those within D behind an element that
passed the test (any superset of them gives the same result, because every element that
remains is a real position with its real L-gram).
"""
from __future__ import annotations

import numpy as np


def lcp32(x: np.ndarray, a: int, b: int) -> int:
    k = 0
    while k < 32 and x[a + k] == x[b + k]:
        k += 1
    return k


def lstar_rank(x: np.ndarray, n: int, W: int, t: int, tile: int | None = None, stats: list | None = None,
               direct_level2: bool = False) -> np.ndarray:
    """x: padded input (n data bytes, then at least W zero bytes).  tile: when given, the
    participant rule is evaluated the way the kernel does it (exact inside a tile of `tile`
    consecutive array entries, conservative at the tile start).  direct_level2: level 2 keeps
    every element (the kernels sort level 2 from x instead of pruning after level 1)."""
    ls = np.zeros(n, dtype=np.uint8)
    D = W - 33 if W > 33 else 0
    if t <= 0 or D == 0 or n == 0:
        return ls
    M = n + D
    pos = np.arange(M, dtype=np.int64)
    key = x[:M].astype(np.int64)
    order = np.argsort(key, kind="stable")
    key, pos = key[order], pos[order]
    for L in range(1, 33):
        m = len(pos)
        if stats is not None:
            stats.append((L, m, int(len(np.unique(key)))))
        if m < t + 2:
            break
        i = np.arange(m)
        j = np.minimum(i + t + 1, m - 1)
        act = (i + t + 1 < m) & (key[j] == key) & (pos[j] - pos <= D)
        if L == 1:
            # first byte occurs <= t times in the window: tc* = count_1 - 1 < t
            for ii in np.nonzero(~act & (pos < n))[0]:
                p = int(pos[ii])
                best, c1 = 32, 0
                k = ii + 1
                while k < m and key[k] == key[ii] and pos[k] <= p + D:
                    c1 += 1
                    best = min(best, lcp32(x, p, int(pos[k])))
                    k += 1
                ls[p] = best if c1 >= 2 else 0
        # only positions that are searched count from here on: an element of the trailing halo
        # matters as a follower of a searched position, never by itself
        act &= pos < n
        ls[pos[act]] = L
        if L == 32 or not act.any():
            break
        # participants of level L+1: within D behind an element that passed (same L-gram)
        if direct_level2 and L == 1:
            part = np.ones(m, dtype=bool)
        elif tile is None:
            last = np.maximum.accumulate(np.where(act, i, -1))
            lc = np.maximum(last, 0)
            part = (last >= 0) & (key[lc] == key) & (pos - pos[lc] <= D)
        else:
            part = np.zeros(m, dtype=bool)
            for s in range(0, m, tile):
                e = min(m, s + tile)
                ii = np.arange(s, e)
                last = np.maximum.accumulate(np.where(act[s:e], ii, -1))
                # carry-in: pretend the element just before the tile passed (superset)
                last = np.where(last < 0, s - 1, last)
                lc = np.maximum(last, 0)
                part[s:e] = (last >= 0) & (key[lc] == key[s:e]) & (pos[s:e] - pos[lc] <= D)
        head = np.ones(m, dtype=bool)
        head[1:] = key[1:] != key[:-1]
        rank = np.cumsum(head) - 1
        nkey = rank[part] * 256 + x[pos[part] + L].astype(np.int64)
        pos = pos[part]
        order = np.argsort(nkey, kind="stable")
        key, pos = nkey[order], pos[order]
    return ls


def merge_pass_order(nkey: np.ndarray) -> np.ndarray:
    """Round-2 design note, stated here so that its exactness is pinned on the CPU: the sort between two
    levels needs ONE ranked radix pass, not three, when the rank range is small enough for a dense count
    matrix.

    nkey[i] = rank[i] * 256 + byte[i] in the order the level kernel writes it: sorted by (rank, position),
    positions increasing inside a rank.  Wanted: the stable order by (rank, byte).

      pass A   stable counting sort by byte alone (the existing 8-bit pass) -> order (byte, rank, position);
               inside it every (byte, rank) pair is one contiguous run, positions still increasing
      matrix   cnt[rank][byte], accumulated where the keys are written; base = its exclusive scan in
               (rank, byte) order, i.e. in key order -- the destination of the first element of each run
      pass B   "merge pass": element i of the pass-A array goes to base[nkey] + (i - start of its run);
               whole runs move as blocks: no ranking (no MATCH.ANY), no per-digit chained prefix

    Returns the permutation (indices into nkey) this produces; equal to np.argsort(nkey, kind='stable')."""
    m = len(nkey)
    if m == 0:
        return np.zeros(0, dtype=np.int64)
    byte = nkey & 255
    a = np.argsort(byte, kind="stable")                      # pass A
    ka = nkey[a]
    groups = int(nkey.max() >> 8) + 1
    cnt = np.bincount(nkey, minlength=groups * 256)          # the count matrix, flattened in key order
    base = np.cumsum(cnt) - cnt
    head = np.ones(m, dtype=bool)
    head[1:] = ka[1:] != ka[:-1]
    start = np.maximum.accumulate(np.where(head, np.arange(m), 0))
    dest = base[ka] + (np.arange(m) - start)                 # pass B
    out = np.empty(m, dtype=np.int64)
    out[dest] = a
    return out
