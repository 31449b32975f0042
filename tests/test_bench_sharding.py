"""bench.py's strong-scaling plumbing on the CPU: ONE input in one shared host buffer, cut into
contiguous position ranges; every rank writes its shard of the table into one shared table.

world_size-2 gloo stands in for the one-rank-per-GPU launch; the oracle stands in for the device
(there is no GPU in this container).  What is under test is bench.shard_cuts (the partition
x3s_search_host uses, x3_search_api.cu), the SharedBuf mappings seen by several processes, the
halo each rank reads out of the shared input and the sha256 of the assembled table."""
import hashlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def test_shard_cuts_cover_align_and_match_the_package():
    pkg = g.load_package()
    for n in (0, 1, 4095, 4096, 10_192_446, 211_938_580):
        for world in (1, 2, 3, 4, 8):
            cuts = bench.shard_cuts(n, world)
            assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == world + 1
            assert all(cuts[i] <= cuts[i + 1] for i in range(world))
            assert all(c % 4096 == 0 for c in cuts[:-1])
            assert [(cuts[r], cuts[r + 1]) for r in range(world)] == pkg.shard_ranges(n, world)


def test_part_pieces_cover_the_input_in_turn():
    """the partition of the N > 1 legs (x3s_search_host_part / x3s_search_device_part): pieces dealt out in turn"""
    for n in (0, 1, 4095, 4096, 10_192_446, 211_938_580):
        for world in (1, 2, 3, 4, 8):
            for piece in (4096, 24592 * 170):
                parts = bench.part_pieces(n, world, piece)
                assert len(parts) == world
                flat = sorted(pc for r in parts for pc in r)
                assert sum(ln for (_, ln) in flat) == n
                pos = 0
                for (p0, ln) in flat:                     # contiguous, no overlap, no gap
                    assert p0 == pos and 0 < ln <= piece and p0 % piece == 0
                    pos += ln
                for r, lst in enumerate(parts):           # rank r: pieces r, r + world, ...
                    assert all((p0 // piece) % world == r for (p0, _) in lst)
                sizes = [sum(ln for (_, ln) in r) for r in parts]
                assert max(sizes) - min(sizes) <= piece


def test_parallel_corpus_generation_is_byte_identical():
    corpus = g.load_submodule("corpus")
    a = corpus.silesia_mix(2_000_000, 5)
    b = corpus.silesia_mix(2_000_000, 5, workers=3)
    assert a == b == corpus.generate("C5", 2_000_000)


def test_generate_cached_round_trips(tmp_path):
    corpus = g.load_submodule("corpus")
    a = corpus.generate_cached("C5", 300_000, workers=2, cache_dir=str(tmp_path))
    b = corpus.generate_cached("C5", 300_000, workers=2, cache_dir=str(tmp_path))
    assert a == b == corpus.generate("C5", 300_000)
    assert (tmp_path / "x3b200_corpus_C5_300000.bin").stat().st_size == 300_000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, W, t, tag, piece):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol
    corpus = g.load_submodule("corpus")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    X = T = None
    if rank == 0:
        X = bench.SharedBuf(tag + "_x", n + W + 4096, True)
        T = bench.SharedBuf(tag + "_l", n, True)
        X.arr[:n] = np.frombuffer(corpus.generate("C5", n), dtype=np.uint8)
        X.arr[n:] = 0
        T.arr[:] = 0xEE
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dist.barrier()
    if rank != 0:
        X = bench.SharedBuf(tag + "_x", n + W + 4096, False)
        T = bench.SharedBuf(tag + "_l", n, False)
    dist.barrier()
    if rank == 0:
        X.unlink()
        T.unlink()
    if piece == 0:
        cuts = bench.shard_cuts(n, world)
        mine = [(cuts[rank], cuts[rank + 1] - cuts[rank])] if cuts[rank + 1] > cuts[rank] else []
    else:
        mine = bench.part_pieces(n, world, piece)[rank]
    for (a0, ln) in mine:
        # what a search of the range sees: its positions followed by W bytes of halo
        a1 = a0 + ln
        sl = np.array(X.arr[a0: a1 + W])
        _, ls = ol.table(sl[: a1 - a0 + W], W, t, p0=0, p1=a1 - a0)
        T.arr[a0:a1] = ls
    dist.barrier()
    if rank == 0:
        _, ref = ol.table(np.array(X.arr[:n]), W, t)
        ok = hashlib.sha256(T.arr[:n]).hexdigest() == hashlib.sha256(ref).hexdigest()
        Path(f"/tmp/{tag}.ok").write_text("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_one_input_sharded_into_one_shared_table_world2_gloo():
    tag = f"pytest_{os.getpid()}"
    mp.spawn(_worker, args=(2, _free_port(), 40_000, 8192, 15, tag, 0), nprocs=2, join=True)
    res = Path(f"/tmp/{tag}.ok")
    assert res.read_text() == "1"
    res.unlink()


def test_one_input_dealt_out_in_pieces_into_one_shared_table_world2_gloo():
    """the N > 1 partition of bench.py: pieces in turn, each read with the window behind it"""
    tag = f"pytest_pc_{os.getpid()}"
    mp.spawn(_worker, args=(2, _free_port(), 40_000, 8192, 15, tag, 6000), nprocs=2, join=True)
    res = Path(f"/tmp/{tag}.ok")
    assert res.read_text() == "1"
    res.unlink()
