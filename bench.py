#!/usr/bin/env python
"""bench.py -- the x3 forward-window match search on B200, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path (histogram loop + threshold selection of
the reference find_best_match, backend.c:58-78) over ONE input: BASELINE.json configs[4], the
211 938 580-byte Silesia-tar-shaped synthetic corpus (C5) with the default flags (-t 15 -w 8) --
the config the metric is quoted on at 1/2/4/8 GPUs.  One position = one input byte = one unit.

N > 1 (launched by torchrun, one rank per GPU): the ONE input lies in one host buffer shared by the
ranks (a tmpfs mapping); it is partitioned into N contiguous position ranges, rank r searches
[a_r, a_r+1) and reads its trailing window halo (W bytes of the next range, or the reference's zero
padding behind the last one: x3.c:579,590; backend.c:60-74 reads p .. p+W-2) straight from that
buffer; every rank's Lstar shard lands in ONE host table.  No data-path collective exists in the
algorithm.  Total work is fixed: "scaling": "strong".  Before anything is timed the sha256 of the
whole table is compared with the oracle-derived hash of the 1-GPU table (tests/golden/tables.json).

  value     positions searched per second (MB/s, 1e6 B): the n positions of C5 / the slowest
            rank's device time per step, shards resident in HBM, CUDA events around the search
  e2e       the same from host memory to host memory through the C ABI a host binds
            (x3s_search_host on each rank's range of the shared buffers: H2D + kernels + D2H inside
            the timed region), max over ranks; `e2e.plugin` is the backend.h drop-in call itself
            (x3_search_prepare on the reference's malloc'ed buffer, all N GPUs from one process,
            measured on rank 0 while the other ranks idle)
  roofline  the dominant kernel against the measured HBM peak
  cpu_baseline / --impl reference
            the UNMODIFIED reference find_best_match (oracle/_ref/libx3ref.so, compiled from
            /root/reference by oracle/Makefile) on all host cores over a bounded sample of the same
            positions.  Nothing under oracle/ is on the product path.
  compress  the first half of BASELINE.json's metric: whole-file `x3 -z` MB/s of the product host
            program with N GPUs (C5 whole), the reference binary timed on a prefix, and the stream
            KAT of the 16 MB scaled C5 (tests/golden/streams_big.json, recorded from the reference)
  c2        secondary object: BASELINE.json configs[1] (10 192 446 B text) through the same legs
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import mmap
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOAD = "C5"
N_BYTES = 211_938_580
W_BYTES = 8192
T_COUNT = 15
METRIC = "match_search_throughput"
UNIT = "MB/s"
WORKLOAD_TEXT = ("C5: 211 938 580 B Silesia-tar-shaped synthetic mix, -t 15 -w 8 (BASELINE.json configs[4]), ONE input "
                 "partitioned by contiguous position ranges with trailing window halo; one position = one byte")
TABLES = ROOT / "tests" / "golden" / "tables.json"
BIG_KATS = ROOT / "tests" / "golden" / "streams_big.json"


def shard_cuts(n: int, world: int, align: int = 4096):
    """Contiguous position ranges of one input over `world` GPUs (the partition x3s_search_host
    itself uses, x3_search_api.cu: cuts at multiples of 4096)."""
    cuts = [min(n, (n * g // world) // align * align) for g in range(world)] + [n]
    return cuts


def part_pieces(n: int, world: int, piece: int):
    """Position ranges of ONE input dealt out to `world` ranks in turn, `piece` positions each (the partition
    x3s_search_host_part / x3s_search_device_part use: x3_search_api.cu, piece = x3s_part_positions(W)): rank r takes
    the pieces r, r + world, ...  Returns one list of (first position, length) per rank."""
    out = [[] for _ in range(world)]
    if piece <= 0:
        return out
    for q in range((n + piece - 1) // piece):
        out[q % world].append((q * piece, min(piece, n - q * piece)))
    return out


# ----------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [ln for (ts, ln) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        for ln in rows:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# CPU arm: the compiled reference (or the oracle port) over a bounded sample
# ----------------------------------------------------------------------------------------
class CpuSearch:
    """find_best_match over sample positions of one padded buffer, on `cores` threads."""

    def __init__(self):
        import oracle_lib as ol  # checker / baseline only -- never on the product path
        self.ol = ol
        self.ora = ol.oracle()
        if ol.have_ref():
            self.kind = "reference"
            R = ol.ref()
            R.set_forward_window(W_BYTES)
            R.set_max_match_count(T_COUNT)
            R.set_magic_factor1(4)
            R.set_magic_factor2(0)
            self.fn = C.cast(R.find_best_match, C.c_void_p)
        else:
            self.kind = "port"
            self.fn = None
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def _one(self, xptr: int, i0: int, i1: int, stride: int):
        if self.fn is not None:
            self.ora.x3o_call_range(self.fn, xptr, i0, i1, stride)
        else:
            # oracle port of backend.c:56-100 with an empty dictionary
            for i in range(i0, i1, stride):
                self.ora.x3o_find_best_match(xptr + i, W_BYTES, T_COUNT, 4, 0, None, None)

    def run(self, xptr: int, n: int, positions: int, threads: int) -> float:
        """Searches `positions` positions spread evenly over [0, n); returns seconds."""
        stride = max(1, n // max(1, positions))
        span = stride * positions
        per = (span // threads // stride + 1) * stride
        thr = []
        t0 = time.perf_counter()
        for k in range(threads):
            i0, i1 = k * per, min(n, (k + 1) * per, span)
            if i0 >= i1:
                continue
            th = threading.Thread(target=self._one, args=(xptr, i0, i1, stride))
            th.start()
            thr.append(th)
        for th in thr:
            th.join()
        return time.perf_counter() - t0

    def calibrate(self, xptr: int, n: int) -> float:
        """positions per second per core (short probe)"""
        probe = 2000 if self.fn is not None else 200
        dt = self.run(xptr, n, probe, 1)
        return probe / dt


def positions_done(n: int, positions: int) -> int:
    stride = max(1, n // max(1, positions))
    span = min(n, stride * positions)
    return (span + stride - 1) // stride


def padded_c5(corpus) -> np.ndarray:
    data = corpus.generate_cached("C5")
    assert len(data) == N_BYTES
    x = np.zeros(N_BYTES + W_BYTES + 64, dtype=np.uint8)
    x[:N_BYTES] = np.frombuffer(data, dtype=np.uint8)
    return x


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    corpus = g.load_submodule("corpus")
    cpu = CpuSearch()
    x = padded_c5(corpus)
    n = N_BYTES
    rate1 = cpu.calibrate(x.ctypes.data, n)
    total_steps = args.steps + args.warmup
    per_step_s = max(0.5, min(6.0, 150.0 / max(1, total_steps)))
    positions = int(max(cpu.cores * 64, min(n, rate1 * cpu.cores * per_step_s)))
    for _ in range(args.warmup):
        cpu.run(x.ctypes.data, n, positions, cpu.cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu.run(x.ctypes.data, n, positions, cpu.cores)
    done = positions_done(n, positions)
    value = done * args.steps / t / 1e6
    sample = (f"{done} of {n} positions per step (every {max(1, n // positions)}th), unmodified reference "
              f"find_best_match, empty dictionary, {cpu.cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT, "window_bytes": W_BYTES, "max_match_count": T_COUNT},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_near_gpu(torch, local):
    """One rank per GPU: run this rank's host threads on the CPUs NVML names as closest to its GPU.
    Returns the CPU set or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:  # NVML missing or the query unsupported: leave the scheduler alone
        pass
    return None


class SharedBuf:
    """One host buffer seen by every rank: a tmpfs file mapped read-write (rank 0 creates it)."""

    def __init__(self, tag: str, nbytes: int, create: bool):
        self.path = f"/dev/shm/x3b200_bench_{tag}"
        self.nbytes = nbytes
        if create:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        self.f = open(self.path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), nbytes)
        self.arr = np.frombuffer(self.mm, dtype=np.uint8)
        self.ptr = self.arr.ctypes.data

    def unlink(self):
        try:
            os.unlink(self.path)
        except OSError:
            pass


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_b200(args):
    import __graft_entry__ as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    corpus = g.load_submodule("corpus")
    tag = f"{os.environ.get('MASTER_PORT', 'solo')}_{os.getppid() if world > 1 else os.getpid()}"
    n = N_BYTES
    # ---- the ONE input, in one host buffer shared by the ranks (before CUDA is initialised:
    # the corpus generator forks workers) --------------------------------------------------
    xbytes = n + W_BYTES + 4096
    c2_n = 10_192_446
    hx = hl = None
    if rank == 0:
        hx = SharedBuf(tag + "_x", xbytes, True)
        hl = SharedBuf(tag + "_l", n, True)
        data = corpus.generate_cached("C5")
        assert len(data) == n
        hx.arr[:n] = np.frombuffer(data, dtype=np.uint8)
        hx.arr[n:] = 0
        del data
        c2 = np.frombuffer(corpus.generate("C2"), dtype=np.uint8)
        hc2x = SharedBuf(tag + "_c2x", c2_n + W_BYTES + 4096, True)
        hc2l = SharedBuf(tag + "_c2l", c2_n, True)
        hc2x.arr[:c2_n] = c2
        hc2x.arr[c2_n:] = 0

    import torch
    pkg = g.load_package()  # raises if lib/libx3b200.so is missing: there is no fallback
    if not torch.cuda.is_available() or pkg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible (the search has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    near = bind_near_gpu(torch, local) if os.environ.get("X3_BENCH_NO_BIND") is None else None
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up; stdout carries
        # exactly one JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=20))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pkg.set_devices([local])
    L = pkg.lib()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if rank != 0:
        hx = SharedBuf(tag + "_x", xbytes, False)
        hl = SharedBuf(tag + "_l", n, False)
        hc2x = SharedBuf(tag + "_c2x", c2_n + W_BYTES + 4096, False)
        hc2l = SharedBuf(tag + "_c2l", c2_n, False)
    barrier()
    if rank == 0:
        for b in (hx, hl, hc2x, hc2l):
            b.unlink()  # the mappings stay; nothing is left behind in tmpfs when the run dies

    golden = json.loads(TABLES.read_text()) if TABLES.exists() else {}

    PIECE = int(L.x3s_part_positions(W_BYTES))

    def legs(name, X, T, nn, steps, warm, flush, check_sha):
        """device-resident leg and host-to-host leg of one input split over the ranks"""
        if world == 1:
            mine = [(0, nn)]
        else:
            mine = part_pieces(nn, world, PIECE)[rank]
        np_r = sum(ln for (_, ln) in mine)
        need = pkg.required_bytes(nn, W_BYTES)
        # page-lock the two shared buffers where they lie (every rank reads its pieces and the windows behind them
        # out of the one input and writes its pieces of the one table)
        reg = []
        for (p, b) in ((X.ptr, nn + W_BYTES), (T.ptr, nn)):
            if L.x3s_host_register(p, b) == 0:
                reg.append(p)
        # (1) input resident in HBM, CUDA events around the search of this rank's pieces (ONE launch)
        d_x = torch.zeros(max(need, 16), dtype=torch.uint8, device=dev)
        d_x[:nn + W_BYTES].copy_(torch.from_numpy(X.arr[:nn + W_BYTES]))
        d_l = torch.zeros(max(nn, 1), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream()

        def step_device():
            if world == 1:
                pkg.search_device(local, d_x.data_ptr(), nn, W_BYTES, T_COUNT, d_l.data_ptr(), None, stream.cuda_stream)
            else:
                rc = L.x3s_search_device_part(local, d_x.data_ptr(), nn, W_BYTES, T_COUNT, d_l.data_ptr(),
                                              stream.cuda_stream, rank, world)
                if rc != 0:
                    raise SystemExit("x3s_search_device_part: " + L.x3s_last_error().decode())

        for _ in range(warm):
            step_device()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for (e0, e1) in ev:
            flush.fill_(1)               # L2 flush between timed iterations, outside the event pair
            e0.record(stream)
            step_device()
            e1.record(stream)
        barrier()
        dev_ms = sum(e0.elapsed_time(e1) for (e0, e1) in ev)
        lstar_dev = d_l.cpu().numpy()
        # (2) host memory to host memory through the C ABI
        tm = pkg.Timing()

        def step_host():
            if world == 1:
                rc = L.x3s_search_host(X.ptr, nn, W_BYTES, T_COUNT, 1, pkg.KERNEL_DEFAULT, T.ptr, None, C.byref(tm))
            else:
                rc = L.x3s_search_host_part(X.ptr, nn, W_BYTES, T_COUNT, T.ptr, C.byref(tm), rank, world)
            if rc != 0:
                raise SystemExit("x3s_search_host(_part): " + L.x3s_last_error().decode())
            return tm.launches

        for _ in range(warm):
            step_host()
        barrier()
        # the whole table, assembled from every rank's pieces in the one host buffer, against the
        # oracle-derived hash of the 1-GPU table -- before anything is timed
        sha = None
        if rank == 0:
            sha = hashlib.sha256(T.arr[:nn]).hexdigest()
            want = golden.get(name, {}).get("lstar_sha256")
            if check_sha and want is not None and sha != want:
                raise SystemExit(f"bench.py: {name} table over {world} GPU(s) has sha256 {sha}, expected {want}")
        for (p0, ln) in mine:
            if not np.array_equal(T.arr[p0:p0 + ln], lstar_dev[p0:p0 + ln]):
                raise SystemExit(f"bench.py: rank {rank}: device-resident and host-buffer runs disagree on {name} at piece {p0}")
        barrier()
        t0 = time.perf_counter()
        launches = 0
        for _ in range(steps):
            launches += step_host()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        for p in reg:
            L.x3s_host_unregister(p)
        per_rank = [dev_ms / steps]
        if dist is not None:
            mine_t = torch.tensor([dev_ms / steps, e2e_s / steps * 1e3, float(launches)], dtype=torch.float64, device=dev)
            every = [torch.zeros_like(mine_t) for _ in range(world)]
            dist.all_gather(every, mine_t)
            per_rank = [float(v[0]) for v in every]
            dev_ms = max(float(v[0]) for v in every) * steps
            e2e_s = max(float(v[1]) for v in every) * steps / 1e3
            launches = int(sum(float(v[2]) for v in every))
        del d_x, d_l
        allp = [(0, nn)] if world == 1 else [pc for r in part_pieces(nn, world, PIECE) for pc in r]
        return {"ms_per_step": dev_ms / steps, "value": nn / (dev_ms / steps * 1e-3) / 1e6,
                "e2e_ms_per_step": e2e_s / steps * 1e3, "e2e_value": nn * steps / e2e_s / 1e6,
                "h2d": int(sum(min(ln + W_BYTES + (128 if world > 1 else 0), nn + W_BYTES - p0) for (p0, ln) in allp)),
                "d2h": int(nn), "launches": launches, "per_rank_ms": per_rank, "sha256": sha,
                "registered": len(reg) == 2, "positions_rank0": np_r if rank == 0 else None, "piece": PIECE,
                "positions_per_rank": [sum(ln for (_, ln) in r) for r in part_pieces(nn, world, PIECE)] if world > 1 else [nn]}

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    steps, warm = args.steps, max(3, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    time.sleep(0.3)
    t_wall0 = time.time()
    main = legs("C5", hx, hl, n, steps, warm, flush, True)
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    c2leg = legs("C2", hc2x, hc2l, c2_n, steps, warm, flush, True)

    # ---- which kernel ran: the production choice for these parameters --------------------------
    kernel_kind = pkg.default_kernel(W_BYTES, T_COUNT, False)
    prof_rank = None
    if rank == 0 and kernel_kind == pkg.KERNEL_RANK:
        # the rank search is a chain of launches: one extra search with an event around every launch
        cuts = shard_cuts(n, world)
        np_r = cuts[1]
        need = pkg.required_bytes(np_r, W_BYTES)
        d_x = torch.zeros(need, dtype=torch.uint8, device=dev)
        have = min(np_r + W_BYTES, n + W_BYTES)
        d_x[:have].copy_(torch.from_numpy(hx.arr[:have]))
        d_l = torch.empty(np_r, dtype=torch.uint8, device=dev)
        os.environ["X3_RANK_PROFILE"] = "1"
        pkg.search_device(local, d_x.data_ptr(), np_r, W_BYTES, T_COUNT, d_l.data_ptr(), None,
                          torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        del os.environ["X3_RANK_PROFILE"]
        prof_rank = pkg.rank_profile(local)
        del d_x, d_l
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    del flush
    torch.cuda.empty_cache()

    # ---- the backend.h plug-in call itself, all N GPUs from this one process -----------------
    plugin = plugin_leg(pkg, hx, hl, n, world, max(2, min(steps, 5)))

    ms_per_step = main["ms_per_step"]
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
    if kernel_kind == pkg.KERNEL_SEG:
        roofline = roofline_seg(main, n, world, peak, peak_src)
        device_launches_per_search = 1
    else:
        roofline = roofline_rank(prof_rank, main, n, world, peak, peak_src)
        device_launches_per_search = main["launches"] // max(1, steps * world)

    # ---- CPU baseline on the box's host cores --------------------------------------------
    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu = CpuSearch()
        rate1 = cpu.calibrate(hx.ptr, n)
        positions = int(min(n, max(cpu.cores * 64, rate1 * cpu.cores * 12.0)))
        dt = cpu.run(hx.ptr, n, positions, cpu.cores)
        done = positions_done(n, positions)
        cpu_baseline = {"value": done / dt / 1e6, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                        "sample": f"{done} of {n} positions (every {max(1, n // positions)}th) of the same input, "
                                  f"find_best_match with an empty dictionary, {dt:.1f} s on {cpu.cores} threads",
                        "per_core_positions_per_s": rate1}

    # ---- whole-file compression through the product CLI beside the reference binary ------
    compress = None
    if not args.no_cpu_baseline and not args.no_compress:
        compress = compress_leg(corpus, hx.arr[:n], world)

    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT,
                   "window_bytes": W_BYTES, "max_match_count": T_COUNT, "positions": n,
                   "positions_per_gpu": main["positions_per_rank"],
                   "piece_positions": PIECE if world > 1 else None,
                   "l2": "flushed between timed steps (256 MiB fill, outside the event pairs)",
                   "sharding": "ONE input; contiguous position ranges (pieces of piece_positions positions, each read with its "
                               "trailing window halo) dealt out to the ranks in turn, no collective; every rank's pieces go "
                               "straight from its GPU to one shared host table" if world > 1 else
                               "one GPU: the whole input",
                   "table_sha256": main["sha256"],
                   "table_sha256_expected": golden.get("C5", {}).get("lstar_sha256"),
                   "table_checked_before_timing": True,
                   "per_rank_ms_per_step": main["per_rank_ms"],
                   "host_buffers_page_locked": main["registered"],
                   "bound_cpus": near if near is None else [near[0], near[-1], len(near)]},
        "e2e": {"value": main["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": main["h2d"],
                "d2h_bytes_per_step": main["d2h"], "ms_per_step": main["e2e_ms_per_step"],
                "api": ("x3s_search_host_part" if world > 1 else "x3s_search_host") + " (include/x3_search.h) on the shared host "
                       "buffers (page-locked in place with x3s_host_register)",
                "plugin": plugin},
        # timed regions only: the host leg's launches (a shard goes piece by piece) + the device leg's
        "gpu_launches": main["launches"] + device_launches_per_search * steps * world,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "compress": compress,
        "c2": {"workload": "C2: 10 192 446 B dickens-shaped text, -t 15 -w 8 (BASELINE.json configs[1]), same legs",
               "value": c2leg["value"], "ms_per_step": c2leg["ms_per_step"], "e2e_value": c2leg["e2e_value"],
               "e2e_ms_per_step": c2leg["e2e_ms_per_step"], "launches_per_search": c2leg["launches"] // max(1, steps),
               "table_sha256": c2leg["sha256"], "table_sha256_expected": golden.get("C2", {}).get("lstar_sha256")},
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    return 0


def roofline_seg(main, n, world, peak, peak_src):
    """Dominant (and only) kernel of the segment search: x3_seg_kernel, ONE launch per device-resident
    search, so its launch duration is the device leg's CUDA-event time per step (events on the launching
    stream around exactly that launch and the 4-byte memset of its segment counter).  Algorithmic bytes
    (production mode): 1 B read + 1 B written per position + the shard's trailing halo.
    The kernel's own HBM traffic is that minimum times (B + D) / B for the re-read halo of every segment
    (profiles/traffic.json, ncu); everything else happens in shared memory, which is what bounds it."""
    np0 = main["positions_per_rank"][0]
    npieces0 = 1 if world == 1 else -(-np0 // max(1, main.get("piece", np0)))
    launch_ms = main["per_rank_ms"][0]
    abytes = 2.0 * np0 + npieces0 * (W_BYTES - 2)     # rank 0's positions: 1 B in, 1 B out, + the window behind each piece
    achieved = abytes / (launch_ms * 1e-3) / 1e9
    ms_per_step = main["ms_per_step"]
    algo_bytes = 2 * n + (1 if world == 1 else -(-n // max(1, main.get("piece", n)))) * (W_BYTES - 2)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "x3_seg_kernel",
                "algorithmic_bytes_per_launch": abytes, "launches_per_search": 1, "avg_launch_ms": launch_ms,
                "share_of_search": 1.0,
                "measured": "CUDA events on the launching stream around rank 0's launch in every timed step of the "
                            "device-resident leg (the search is this one launch)",
                "note": "HBM traffic is at its algorithmic minimum (input once per segment incl. window halo, Lstar once); "
                        "the kernel is bound on chip -- shared-memory wavefronts and issue slots (see on_chip) -- so the "
                        "HBM fraction is small by construction",
                "search": {"algorithmic_bytes": algo_bytes, "achieved_GBps": algo_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac": algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                           "equivalent_pair_tests_per_s": n * (W_BYTES - 33) / (ms_per_step * 1e-3)}}
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            if pj.get("kernel") == "x3_seg_kernel":
                roofline["traffic"] = pj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = pj.get("source")
                roofline["on_chip"] = pj.get("on_chip")
        except (ValueError, OSError):
            pass
    return roofline


def roofline_rank(prof_rank, main, n, world, peak, peak_src):
    """Dominant kernel of the rank search: the radix pass.  One element = one 4-byte key + one
    4-byte position word, read once and written once per pass: 16 algorithmic bytes."""
    radix_ms, radix_el, radix_n = prof_rank["radix"]
    level_ms, level_el, level_n = prof_rank["level"]
    prof_total_ms = sum(v[0] for v in prof_rank.values())
    rbytes = 16.0 * radix_el / max(radix_n, 1)
    achieved = 16.0 * radix_el / (radix_ms * 1e-3) / 1e9 if radix_ms > 0 else 0.0
    ms_per_step = main["ms_per_step"]
    algo_bytes = 2 * n + world * (W_BYTES - 2)      # the whole search: 1 B read + 1 B written per position + halos
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "x3_rank_radix_kernel",
                "algorithmic_bytes_per_launch": rbytes, "launches_per_search": radix_n,
                "avg_launch_ms": radix_ms / max(radix_n, 1), "share_of_search": radix_ms / max(prof_total_ms, 1e-9),
                "measured": "CUDA events around every launch of one extra search of rank 0's shard "
                            "(X3_RANK_PROFILE=1, one lane), elements from the level sizes the device recorded",
                "level_kernel": {"kernel": "x3_rank_level_kernel", "launches_per_search": level_n,
                                 "ms_per_search": level_ms, "elements": level_el,
                                 "achieved_GBps_at_18B": 18.0 * level_el / (level_ms * 1e-3) / 1e9 if level_ms > 0 else None},
                "search": {"algorithmic_bytes": algo_bytes, "achieved_GBps": algo_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac": algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                           "equivalent_pair_tests_per_s": n * (W_BYTES - 33) / (ms_per_step * 1e-3)}}
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            roofline["traffic"] = pj.get("dram_bytes_per_launch")
            roofline["traffic_source"] = pj.get("source")
        except (ValueError, OSError):
            pass
    return roofline


def plugin_leg(pkg, hx, hl, n, world, reps):
    """x3_search_prepare() -- the one call a reference-side maintainer adds behind fload()
    (x3.c:591) -- on a malloc'ed copy of the input the way the reference's
    main() holds it (x3.c:579: pageable memory), with all `world` GPUs driven from this process,
    followed by a find_best_match() sweep over a sample of positions (empty dictionary) and the
    table compared with the sharded legs' table."""
    L = pkg.lib()
    pkg.set_devices([])
    os.environ["X3_GPUS"] = str(world)
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    buf = libc.malloc(n + W_BYTES)
    if not buf:
        return {"unavailable": "malloc failed"}
    C.memmove(buf, hx.ptr, n + W_BYTES)
    L.set_forward_window(W_BYTES)
    L.set_max_match_count(T_COUNT)
    L.set_magic_factor1(4)
    L.set_magic_factor2(0)
    # an empty dictionary (the filter of backend.c:79-90 then never fires), as in the reference arm
    no_find = pkg.DICT_FIND_FN(lambda p: (1 << 64) - 1)
    no_len = pkg.DICT_LEN_FN(lambda i: 0)
    L.x3_backend_set_dict(C.cast(no_find, C.c_void_p), C.cast(no_len, C.c_void_p))
    times = []
    first = []
    same = None
    sweep_s = None
    for r in range(reps + 1):
        # prepare starts the search on its own thread and returns; the table lands piece by piece from the
        # left (find_best_match(p) waits for p's piece).  Timed: until the first piece has landed (when the
        # reference's left-to-right compress() can start) and until the whole table has (x3_search_wait)
        t0 = time.perf_counter()
        L.x3_search_prepare(buf, n)
        while L.x3_search_ready() == 0:
            pass
        t_first = time.perf_counter() - t0
        L.x3_search_wait()
        dt = time.perf_counter() - t0
        if r > 0:
            times.append(dt)
            first.append(t_first)
        if r == reps:
            Hp, Lp, nn = C.c_void_p(), C.c_void_p(), C.c_size_t()
            L.x3_search_table(C.byref(Hp), C.byref(Lp), C.byref(nn))
            tab = np.ctypeslib.as_array(C.cast(Lp, C.POINTER(C.c_uint8)), shape=(nn.value,))
            same = bool(nn.value == n and np.array_equal(tab, hl.arr[:n]))
            stride = max(1, n // 2_000_000)
            t1 = time.perf_counter()
            acc = 0
            fbm = L.find_best_match
            for p in range(0, n, stride * 64):  # python call overhead dominates: a thin sample only
                acc += fbm(buf + p)
            sweep_s = time.perf_counter() - t1
        L.x3_search_release()
    L.x3_backend_set_dict(None, None)
    libc.free(buf)
    pkg.set_devices([])
    best = min(times)
    return {"value": n / best / 1e6, "unit": UNIT, "ms_per_call": best * 1e3, "ms_per_call_all": [t * 1e3 for t in times],
            "ms_until_first_piece": min(first) * 1e3, "gpus": world, "table_equals_sharded_legs": same,
            "api": "x3_search_prepare(iptr, isize) + x3_search_wait() (include/x3_backend.h) on malloc'ed (pageable) memory, "
                   "one process, X3_GPUS=N; the table is fresh pageable memory as well",
            "find_best_match_sweep_s": sweep_s}


def _elapsed(stderr: str, key: str = "elapsed time:"):
    for ln in stderr.splitlines():
        if ln.startswith(key):
            return float(ln.split(":")[1])
    return None


def compress_leg(corpus, c5: np.ndarray, world: int):
    """x3 -z end to end (file in, .x3 out), the first half of BASELINE.json's metric: the product
    binary (GPU search on `world` GPUs + re-designed sequential pass) on the whole C5 file; the same
    binary on the 16 MB scaled C5 whose reference stream is pinned in tests/golden/streams_big.json
    (byte-identity without re-running 3 minutes of reference); and the unmodified reference binary
    timed on the first 1 MB of C5 (its time is ~20 s per MB, single-threaded by design)."""
    import tempfile
    x3 = ROOT / "x3-compressor_b200" / "bin" / "x3"
    ref = ROOT / "oracle" / "_ref" / "x3_ref"
    if not x3.exists():
        return {"unavailable": "x3-compressor_b200/bin/x3 not built"}
    env = dict(os.environ, X3_GPUS=str(world))
    out = {}
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=tmpdir) as td:
        td = Path(td)
        (td / "c5").write_bytes(c5.tobytes())
        t0 = time.perf_counter()
        try:
            r = subprocess.run([str(x3), "-zf", str(td / "c5"), str(td / "c5.x3")], stderr=subprocess.PIPE, text=True,
                               env=env, timeout=300)
        except subprocess.TimeoutExpired:
            return {"unavailable": "bin/x3 -z on C5 did not finish in 300 s"}
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"unavailable": "bin/x3 failed: " + r.stderr[-200:]}
        el = _elapsed(r.stderr)
        srch = _elapsed(r.stderr, "of which match search")
        start = _elapsed(r.stderr, "of which CUDA start-up") or 0.0
        stream_len = (td / "c5.x3").stat().st_size
        ok = None
        dec = None
        try:
            rd = subprocess.run([str(x3), "-df", str(td / "c5.x3"), str(td / "c5.back")], stderr=subprocess.PIPE,
                                text=True, timeout=300)
            ok = rd.returncode == 0 and (td / "c5.back").read_bytes() == c5.tobytes()
            dec = len(c5) / _elapsed(rd.stderr) / 1e6 if ok else None
        except subprocess.TimeoutExpired:
            ok = None
        out.update({"value": len(c5) / el / 1e6, "unit": "MB/s", "bytes": len(c5), "gpus": world, "elapsed_s": el,
                    "search_s": srch, "cuda_startup_s": start,
                    "value_excluding_cuda_startup": len(c5) / max(el - start, 1e-9) / 1e6,
                    "host_pass_s": el - (srch or 0.0), "process_wall_s": wall, "stream_bytes": stream_len,
                    "ratio": len(c5) / stream_len, "round_trip_ok": ok, "decompress_MB_per_s": dec,
                    "what": "bin/x3 -z on the whole C5 file: GPU search (incl. transfers) + sequential host pass, the "
                            "program's own 'elapsed time' (brackets prepare + compress, cf. x3.c:597-601)"})
        kats = json.loads(BIG_KATS.read_text()) if BIG_KATS.exists() else {}
        kat = kats.get("C5:16000000:")
        if kat is not None:
            small = corpus.generate("C5", 16_000_000)
            (td / "c5s").write_bytes(small)
            rs = subprocess.run([str(x3), "-zf", str(td / "c5s"), str(td / "c5s.x3")], stderr=subprocess.PIPE, text=True, env=env)
            s = (td / "c5s.x3").read_bytes() if rs.returncode == 0 else b""
            out["kat_16MB_scaled_C5"] = {"stream_identical_to_reference": hashlib.sha256(s).hexdigest() == kat["sha256"]
                                         and len(s) == kat["len"], "sha256": hashlib.sha256(s).hexdigest(),
                                         "ours_MB_per_s": 16.0 / (_elapsed(rs.stderr) or float("nan")),
                                         "reference_MB_per_s_build_container": 16.0 / kat["ref_elapsed_s"] if kat.get("ref_elapsed_s") else None}
        if ref.exists():
            npre = 1_000_000
            (td / "pre").write_bytes(c5[:npre].tobytes())
            rr = subprocess.run([str(ref), "-zf", str(td / "pre"), str(td / "pre.ref.x3")], stderr=subprocess.PIPE, text=True)
            ro = subprocess.run([str(x3), "-zf", str(td / "pre"), str(td / "pre.x3")], stderr=subprocess.PIPE, text=True, env=env)
            same = (rr.returncode == 0 and ro.returncode == 0 and
                    (td / "pre.ref.x3").read_bytes() == (td / "pre.x3").read_bytes())
            rel = _elapsed(rr.stderr)
            out["reference"] = {"value": npre / rel / 1e6, "unit": "MB/s", "cores": 1, "elapsed_s": rel,
                                "sample": f"first {npre} B of C5, unmodified reference x3 -z (single-threaded by design)",
                                "ours_on_same_prefix_MB_per_s": npre / _elapsed(ro.stderr) / 1e6,
                                "stream_identical_on_prefix": same}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-compress", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
