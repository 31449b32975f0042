"""How should the per-GPU Lstar shards reach the host pass?  (BASELINE north_star:
"NVLink P2P copy ... NCCL only if measured faster".)  One process, G GPUs, one shard of
`n` bytes per GPU; the consumer is host code, so the end point is pinned host memory.

  direct   every GPU copies its shard straight to pinned host memory (its own PCIe link)
  p2p      every GPU's shard is copied to GPU 0 over NVLink, then GPU 0 copies all of it to the host

    python tests/gpu_gather_bench.py [bytes_per_gpu]
"""
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_192_446
G = torch.cuda.device_count()
shards = [torch.randint(0, 33, (n,), dtype=torch.uint8, device=f"cuda:{g}") for g in range(G)]
host = torch.empty(G * n, dtype=torch.uint8).pin_memory()
streams = [torch.cuda.Stream(device=f"cuda:{g}") for g in range(G)]
gather0 = torch.empty(G * n, dtype=torch.uint8, device="cuda:0")
for g in range(1, G):
    torch.cuda.set_device(0)
    try:
        torch.cuda.can_device_access_peer(0, g)
    except Exception:
        pass


def sync_all():
    for g in range(G):
        torch.cuda.synchronize(g)


def direct():
    for g in range(G):
        with torch.cuda.stream(streams[g]):
            host[g * n:(g + 1) * n].copy_(shards[g], non_blocking=True)
    sync_all()


def p2p():
    for g in range(G):
        with torch.cuda.stream(streams[g]):
            gather0[g * n:(g + 1) * n].copy_(shards[g], non_blocking=True)
    sync_all()
    with torch.cuda.stream(streams[0]):
        host.copy_(gather0, non_blocking=True)
    sync_all()


for name, fn in (("direct", direct), ("p2p", p2p)):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:7s} {G} GPUs x {n} B: {dt * 1e3:.3f} ms per gather -> {G * n / dt / 1e9:.1f} GB/s to the host", flush=True)
