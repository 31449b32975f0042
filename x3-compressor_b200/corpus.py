"""Seeded synthetic corpora shaped like the Silesia members.

The Silesia corpus is not available offline, so every benchmark/parity input is
generated here, deterministically (numpy PCG64, fixed seeds, fixed sizes).  The
generators are vectorised so that the 212 MB mix is produced in well under a
minute.  None of this is on the product path; it only feeds tests and bench.py.

Configs (BASELINE.json `configs`, sizes from BASELINE.md section 3):
  C1  1 000 000 B   English-like text, default flags
  C2  10 192 446 B  dickens-shaped text, default flags          (bench headline)
  C3  50 000 000 B  webster/xml-shaped, -w 1024 -t 64
  C4  8 474 240 B   x-ray/sao-shaped low-redundancy binary
  C5  211 938 580 B Silesia-tar-shaped mix
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_W = np.array(
    [12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5,
     1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_LETTER_W = _LETTER_W / _LETTER_W.sum()


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _vocabulary(rng: np.random.Generator, size: int, maxlen: int = 14):
    lens = np.clip(rng.poisson(4.2, size) + 1, 1, maxlen).astype(np.int64)
    # the most frequent words are short, as in natural text
    lens[: min(size, 40)] = np.clip(rng.integers(1, 5, min(size, 40)), 1, maxlen)
    mat = _LETTERS[rng.choice(len(_LETTERS), size=(size, maxlen), p=_LETTER_W)]
    return mat, lens


def _scatter_words(mat, lens, ids, seps, caps):
    """Lay out word[ids[i]] followed by separator seps[i] (0..2 bytes, see below)."""
    sep_bytes = np.array([[32, 0], [44, 32], [46, 32], [10, 0], [46, 10], [59, 32]], dtype=np.uint8)
    sep_len = np.array([1, 2, 2, 1, 2, 2], dtype=np.int64)
    wl = lens[ids]
    sl = sep_len[seps]
    tot = wl + sl
    off = np.concatenate(([0], np.cumsum(tot)[:-1]))
    out = np.empty(int(tot.sum()), dtype=np.uint8)
    for k in range(mat.shape[1]):
        m = wl > k
        if not m.any():
            break
        out[off[m] + k] = mat[ids[m], k]
    for k in range(2):
        m = sl > k
        out[off[m] + wl[m] + k] = sep_bytes[seps[m], k]
    out[off[caps]] -= 32  # capitalise sentence starts
    return out


def text(n: int, seed: int, vocab: int = 5000, zipf_a: float = 1.07, para: bool = False) -> bytes:
    """English-like text: Zipf-weighted pseudo-words, sentences of 5-20 words."""
    rng = _rng(seed)
    mat, lens = _vocabulary(rng, vocab)
    w = 1.0 / np.arange(1, vocab + 1) ** zipf_a
    w /= w.sum()
    nwords = int(n / 4.5) + 64
    ids = rng.choice(vocab, size=nwords, p=w)
    # sentence structure: a period every 5..20 words, commas inside, newlines
    seps = np.zeros(nwords, dtype=np.int64)
    gaps = rng.integers(5, 21, size=nwords // 5 + 2)
    ends = np.cumsum(gaps)
    ends = ends[ends < nwords]
    seps[rng.random(nwords) < 0.07] = 1            # ", "
    seps[rng.random(nwords) < 0.01] = 5            # "; "
    seps[ends] = 2                                 # ". "
    nl = ends[rng.random(len(ends)) < (0.12 if para else 0.3)]
    seps[nl] = 4                                   # ".\n"
    caps = np.zeros(nwords, dtype=bool)
    caps[0] = True
    caps[np.minimum(ends + 1, nwords - 1)] = True
    out = _scatter_words(mat, lens, ids, seps, caps)
    while len(out) < n:                            # pragma: no cover (size estimate is generous)
        out = np.concatenate((out, out))
    return out[:n].tobytes()


def webster(n: int, seed: int) -> bytes:
    """Dictionary-entry-shaped text: HEADWORD, part of speech, definitions."""
    rng = _rng(seed)
    body = np.frombuffer(text(n, seed + 1000, vocab=20000), dtype=np.uint8).copy()
    # every ~60 bytes of body, start a new entry with an upper-case headword line
    pos_tags = [b"n.", b"v. t.", b"a.", b"adv.", b"v. i.", b"prep."]
    out = bytearray()
    i = 0
    hw_mat, hw_len = _vocabulary(rng, 4096)
    cuts = np.cumsum(rng.integers(80, 400, size=n // 200 + 8))
    for c in cuts:
        if len(out) >= n:
            break
        h = int(rng.integers(0, 4096))
        head = bytes(hw_mat[h, : hw_len[h]] - 32)
        tag = pos_tags[int(rng.integers(0, len(pos_tags)))]
        out += b"\n" + head + b"\n" + head.capitalize() + b", " + tag + b" [Etym: L. " + head.lower()[:4] + b"us.]\n"
        k = int(rng.integers(1, 4))
        seg = body[i:c].tobytes()
        i = int(c)
        for d in range(k):
            part = seg[d * len(seg) // k:(d + 1) * len(seg) // k]
            out += str(d + 1).encode() + b". " + part.strip() + b"\n"
    out += body[: max(0, n - len(out))].tobytes()
    return bytes(out[:n])


def xml(n: int, seed: int) -> bytes:
    """XML-like: nested tags from a small tag set, attributes, numeric/text payloads."""
    rng = _rng(seed)
    tags = [b"record", b"item", b"name", b"value", b"entry", b"title", b"author", b"date", b"ref", b"note"]
    words = text(max(n // 3, 4096), seed + 7, vocab=3000).split()
    out = bytearray(b'<?xml version="1.0" encoding="UTF-8"?>\n<root>\n')
    wi = 0
    rid = 0
    while len(out) < n:
        rid += 1
        t0 = tags[int(rng.integers(0, 3))]
        out += b'  <' + t0 + b' id="' + str(rid).encode() + b'" type="' + tags[int(rng.integers(3, 10))] + b'">\n'
        for _ in range(int(rng.integers(2, 7))):
            t1 = tags[int(rng.integers(2, 10))]
            if rng.random() < 0.4:
                payload = str(int(rng.integers(0, 100000))).encode()
            else:
                k = int(rng.integers(1, 6))
                payload = b" ".join(words[wi % len(words): wi % len(words) + k])
                wi += k
            out += b'    <' + t1 + b'>' + payload + b'</' + t1 + b'>\n'
        out += b'  </' + t0 + b'>\n'
    return bytes(out[:n])


def image16(n: int, seed: int, step: int = 40, bits: int = 12) -> bytes:
    """16-bit little-endian random-walk samples (x-ray / mr shaped)."""
    rng = _rng(seed)
    m = n // 2 + 1
    steps = rng.integers(-step, step + 1, size=m)
    walk = np.cumsum(steps)
    hi = (1 << bits) - 1
    # reflect into [0, hi]
    walk = np.abs(((walk + hi) % (2 * hi)) - hi).astype(np.uint16)
    return walk.astype("<u2").tobytes()[:n]


def records(n: int, seed: int, width: int = 28) -> bytes:
    """Fixed-width records of slowly varying fields with noisy low bytes (sao / osdb shaped)."""
    rng = _rng(seed)
    m = n // width + 1
    rec = np.zeros((m, width), dtype=np.uint8)
    base = np.cumsum(rng.integers(0, 3, size=m)).astype(np.uint64)
    for f in range(0, width, 4):
        v = (base * np.uint64(f + 3) + rng.integers(0, 1 << 10, size=m).astype(np.uint64)).astype(np.uint32)
        b = v.view(np.uint8).reshape(m, 4)
        rec[:, f:f + 4] = b[:, : min(4, width - f)]
    rec[:, width - 2:] = rng.integers(0, 4, size=(m, 2), dtype=np.uint8)
    return rec.tobytes()[:n]


def executable(n: int, seed: int) -> bytes:
    """Executable-like byte soup: opcode patterns, zero runs, tables, strings."""
    rng = _rng(seed)
    out = np.empty(n + 4096, dtype=np.uint8)
    pats = [rng.integers(0, 256, size=int(rng.integers(2, 9)), dtype=np.uint8) for _ in range(256)]
    strings = np.frombuffer(text(max(n // 16, 4096), seed + 3, vocab=2000), dtype=np.uint8)
    i = 0
    si = 0
    kinds = rng.integers(0, 100, size=n // 24 + 16)
    k = 0
    while i < n:
        kind = kinds[k % len(kinds)]
        k += 1
        if kind < 55:  # code: a few opcode patterns with random immediates
            for _ in range(int(rng.integers(2, 8))):
                p = pats[int(rng.integers(0, 256) if rng.random() < 0.3 else rng.integers(0, 24))]
                out[i:i + len(p)] = p
                i += len(p)
                imm = int(rng.integers(0, 3))
                out[i:i + imm] = rng.integers(0, 256, size=imm, dtype=np.uint8)
                i += imm
        elif kind < 70:  # zero / fill run
            ln = int(rng.integers(4, 64))
            out[i:i + ln] = 0 if rng.random() < 0.8 else 0xFF
            i += ln
        elif kind < 85:  # table of 32-bit little-endian offsets
            ln = int(rng.integers(4, 24))
            base = int(rng.integers(0, 1 << 20))
            tab = (base + np.cumsum(rng.integers(0, 64, size=ln))).astype("<u4")
            b = tab.view(np.uint8)
            out[i:i + len(b)] = b
            i += len(b)
        elif kind < 95:  # ascii strings
            ln = int(rng.integers(8, 80))
            s = strings[si % (len(strings) - 128): si % (len(strings) - 128) + ln]
            si += ln
            out[i:i + len(s)] = s
            i += len(s)
            out[i] = 0
            i += 1
        else:  # high-entropy blob
            ln = int(rng.integers(16, 256))
            out[i:i + ln] = rng.integers(0, 256, size=ln, dtype=np.uint8)
            i += ln
    return out[:n].tobytes()


def chemdb(n: int, seed: int) -> bytes:
    """Chemical-database-shaped text (nci): numeric tables with long repeats.
    Noise is injected so the overall x3 ratio stays far below the reference
    decoder's 64:1 output-buffer limit (reference x3.c:621)."""
    rng = _rng(seed)
    out = bytearray()
    mol = 0
    while len(out) < n:
        mol += 1
        na = int(rng.integers(8, 40))
        out += f"{mol}\n  -OEChem-0{int(rng.integers(1000000, 9999999))}\n\n".encode()
        out += f"{na:3d}{na - 1:3d}  0     0  0  0  0  0  0999 V2000\n".encode()
        xyz = rng.normal(0, 3, size=(na, 3))
        el = rng.choice([b"C", b"C", b"C", b"H", b"H", b"N", b"O", b"S"], size=na)
        for a in range(na):
            out += f"{xyz[a, 0]:10.4f}{xyz[a, 1]:10.4f}{xyz[a, 2]:10.4f} ".encode() + el[a] + b"   0  0  0  0  0  0  0  0  0  0  0  0\n"
        for a in range(1, na):
            out += f"{int(rng.integers(1, a + 1)):3d}{a + 1:3d}{int(rng.integers(1, 3)):3d}  0  0  0  0\n".encode()
        out += b"M  END\n$$$$\n"
    return bytes(out[:n])


_SILESIA = [
    ("dickens", 10192446, "text"), ("mozilla", 51220480, "exe"), ("mr", 9970564, "img"),
    ("nci", 33553445, "chem"), ("ooffice", 6152192, "exe"), ("osdb", 10085684, "rec"),
    ("reymont", 6627202, "text"), ("samba", 21606400, "mix"), ("sao", 7251944, "rec"),
    ("webster", 41458703, "web"), ("xml", 5345280, "xml"), ("x-ray", 8474240, "img"),
]


def _member(kind: str, n: int, seed: int) -> bytes:
    if kind == "text":
        return text(n, seed, vocab=30000, para=True)
    if kind == "exe":
        return executable(n, seed)
    if kind == "img":
        return image16(n, seed)
    if kind == "chem":
        return chemdb(n, seed)
    if kind == "rec":
        return records(n, seed)
    if kind == "web":
        return webster(n, seed)
    if kind == "xml":
        return xml(n, seed)
    if kind == "mix":
        h = n // 2
        return text(h, seed, vocab=8000) + executable(n - h, seed + 1)
    raise ValueError(kind)


def _member_job(job):
    kind, body, seed = job
    return _member(kind, body, seed)


def silesia_mix(n: int = 211938580, seed: int = 5, workers: int = 1) -> bytes:
    """Tar-like concatenation with the Silesia size profile, scaled to n bytes.  workers > 1
    generates the members in that many processes (same bytes: every member has a seed of its own)."""
    total = sum(s for _, s, _ in _SILESIA)
    heads = []
    jobs = []
    used = 0
    for i, (name, size, kind) in enumerate(_SILESIA):
        share = n - used if i == len(_SILESIA) - 1 else int(size * n / total)
        share = max(share, 0)
        hdr = (name.encode().ljust(100, b"\0") + b"0000644\0" + f"{share:011o}\0".encode()).ljust(512, b"\0")
        body = max(share - 512, 0)
        heads.append(hdr[: min(512, share)])
        jobs.append((kind, body, seed * 100 + i))
        used += share
    if workers > 1:
        import multiprocessing as mp
        # largest members first, so that the pool's wall time is that of the largest one
        order = sorted(range(len(jobs)), key=lambda j: -jobs[j][1])
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            res = pool.map(_member_job, [jobs[j] for j in order], chunksize=1)
        bodies = [b""] * len(jobs)
        for j, b in zip(order, res):
            bodies[j] = b
    else:
        bodies = [_member(*job) if job[1] else b"" for job in jobs]
    out = b"".join(h + (b if job[1] else b"") for h, b, job in zip(heads, bodies, jobs))
    assert len(out) == n, (len(out), n)
    return out


CONFIGS = {
    "C1": dict(size=1_000_000, flags=dict(t=15, w_kb=8), desc="1 MB English-like text, default -t/-w"),
    "C2": dict(size=10_192_446, flags=dict(t=15, w_kb=8), desc="10 MB dickens-shaped text, default -t/-w"),
    "C3": dict(size=50_000_000, flags=dict(t=64, w_kb=1024), desc="50 MB webster/xml-shaped, -w 1024 -t 64"),
    "C4": dict(size=8_474_240, flags=dict(t=15, w_kb=8), desc="8.5 MB x-ray/sao-shaped binary, default"),
    "C5": dict(size=211_938_580, flags=dict(t=15, w_kb=8), desc="211 MB Silesia-tar-shaped mix, default"),
}


def generate(name: str, size: int | None = None) -> bytes:
    """Returns the first `size` bytes of config `name` (prefixes are stable for
    C1/C2/C4; C3/C5 are generated at the requested size)."""
    cfg = CONFIGS[name]
    n = cfg["size"] if size is None else size
    if name == "C1":
        return text(cfg["size"], seed=1, vocab=5000)[:n]
    if name == "C2":
        return text(cfg["size"] if n > 2_000_000 else max(n, 1 << 16), seed=2, vocab=30000, para=True)[:n]
    if name == "C3":
        a = int(n * 0.83)
        b = int(n * 0.106)
        return (webster(a, seed=3) + xml(b, seed=33) + text(n - a - b, seed=34, vocab=12000))[:n]
    if name == "C4":
        a = int(n * 0.6) & ~1
        return (image16(a, seed=4) + records(n - a, seed=44))[:n]
    if name == "C5":
        return silesia_mix(n, seed=5)
    raise KeyError(name)


def generate_cached(name: str, size: int | None = None, workers: int = 8, cache_dir: str = "/dev/shm") -> bytes:
    """generate() through a file cache in tmpfs (the 212 MB mix takes ~40 s of Python; the bench's
    two arms and four GPU counts would each pay it).  The cache is only ever a copy of what
    generate() returns: it is written under a temporary name and renamed, and a file of the wrong
    length is ignored.  Call it before CUDA is initialised (the workers are forked)."""
    import os
    n = CONFIGS[name]["size"] if size is None else size
    path = os.path.join(cache_dir, f"x3b200_corpus_{name}_{n}.bin")
    try:
        if os.path.getsize(path) == n:
            with open(path, "rb") as f:
                return f.read()
    except OSError:
        pass
    data = silesia_mix(n, seed=5, workers=workers) if name == "C5" else generate(name, size)
    try:
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
    except OSError:
        pass
    return data
