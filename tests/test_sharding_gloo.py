"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank takes its contiguous position
range with the trailing window halo (package shard_ranges()), computes its
rows with the oracle (the checker stands in for the device here -- there is no GPU in this
container), rank 0 gathers and compares with the unsharded table.  What is under test is
the partitioning, the halo arithmetic and the gather order used by bench.py / the C ABI."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, W, t, out_path):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import __graft_entry__ as g
    import oracle_lib as ol
    pkg = g.load_package()
    corpus = g.load_submodule("corpus")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.frombuffer(corpus.generate("C5", n), dtype=np.uint8)
    x = np.zeros(n + W, dtype=np.uint8)
    x[:n] = data
    a, b = pkg.shard_ranges(n, world, align=4096)[rank]
    # the bytes this rank needs: its positions + a trailing halo of W bytes
    need = x[a: min(len(x), b + W)]
    sl = np.zeros((b - a) + W, dtype=np.uint8)
    sl[: len(need)] = need
    _, ls = ol.table(sl[: (b - a) + W], W, t, p0=0, p1=b - a) if b > a else (None, np.zeros(0, np.uint8))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([b - a], dtype=torch.int64))
    mx = int(max(s.item() for s in sizes))
    mine = torch.zeros(mx, dtype=torch.uint8)
    mine[: b - a] = torch.from_numpy(ls.copy())
    gathered = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank == 0:
        full = np.concatenate([gathered[r][: int(sizes[r].item())].numpy() for r in range(world)])
        _, ref = ol.table(data, W, t)
        np.save(out_path, np.array([int(np.array_equal(full, ref)), len(full), n]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,W,t", [(30000, 8192, 15), (9000, 1024, 3)])
def test_halo_sharding_world2_gloo(tmp_path, n, W, t):
    out = tmp_path / "res.npy"
    mp.spawn(_worker, args=(2, _free_port(), n, W, t, str(out)), nprocs=2, join=True)
    ok, got, want = np.load(out)
    assert (ok, got) == (1, want)


def test_shard_ranges_cover_and_align(pkg):
    for n in (0, 1, 4095, 4096, 10_192_446, 211_938_580):
        for world in (1, 2, 4, 8):
            r = pkg.shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert all(a % 4096 == 0 for a, _ in r)
