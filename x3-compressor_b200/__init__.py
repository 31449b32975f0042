"""x3-compressor_b200 -- host-side mirror of the C ABI in include/*.h.

The product is the shared library lib/libx3b200.so (hand-written sm_100a kernels
behind `extern "C"` entry points) and the C99 host code around it.  This module
is the thin ctypes binding used by tests/ and bench.py; it mirrors the reference's
backend.h interface (same names, argument meaning and error behaviour, reference
backend.h:20-31) plus the device-level calls of include/x3_search.h.

There is deliberately no fallback: if the library is missing, or no CUDA device
is visible, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libx3b200.so"

X3S_OK = 0
X3S_ERR_CUDA = -1
X3S_ERR_ARG = -2
X3S_ERR_UNSUPP = -3
MAX_MATCH_LEN = 32
KERNEL_DEFAULT, KERNEL_NAIVE, KERNEL_BITSLICED, KERNEL_STREAM, KERNEL_STREAM_FULL, KERNEL_RANK, KERNEL_SEG = 0, 1, 2, 3, 4, 5, 6


class X3SearchError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"x3 search error {code}: {msg}")
        self.code = code


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("total_ms", C.c_double), ("gpus", C.c_int), ("launches", C.c_int)]


DICT_FIND_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p)
DICT_LEN_FN = C.CFUNCTYPE(C.c_size_t, C.c_size_t)

_lib = None


def lib() -> C.CDLL:
    """Loads lib/libx3b200.so (built by `make -C x3-compressor_b200/csrc` or
    __graft_entry__.build()).  Raises if it is not there: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH), mode=os.RTLD_LAZY | os.RTLD_GLOBAL)
    L.x3s_device_count.restype = C.c_int
    L.x3s_last_error.restype = C.c_char_p
    L.x3s_version.restype = C.c_char_p
    L.x3s_required_bytes.restype = C.c_size_t
    L.x3s_required_bytes.argtypes = [C.c_size_t, C.c_size_t]
    L.x3s_search_device.restype = C.c_int
    L.x3s_search_device.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int]
    L.x3s_search_host.restype = C.c_int
    L.x3s_search_host.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.POINTER(Timing)]
    L.x3s_host_alloc.restype = C.c_void_p
    L.x3s_host_alloc.argtypes = [C.c_size_t]
    L.x3s_host_free.argtypes = [C.c_void_p]
    L.x3s_host_register.restype = C.c_int
    L.x3s_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.x3s_host_unregister.restype = C.c_int
    L.x3s_host_unregister.argtypes = [C.c_void_p]
    L.x3s_part_positions.restype = C.c_size_t
    L.x3s_part_positions.argtypes = [C.c_size_t]
    L.x3s_search_device_part.restype = C.c_int
    L.x3s_search_device_part.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int]
    L.x3s_search_host_part.restype = C.c_int
    L.x3s_search_host_part.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(Timing),
                                       C.c_int, C.c_int]
    L.x3s_default_kernel.restype = C.c_int
    L.x3s_default_kernel.argtypes = [C.c_size_t, C.c_int, C.c_int]
    L.x3s_rank_profile.restype = C.c_int
    L.x3s_rank_profile.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.x3s_rank_plan.restype = C.c_int
    L.x3s_rank_plan.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                C.POINTER(C.c_int)]
    L.x3s_release.restype = None
    L.x3s_set_devices.restype = C.c_int
    L.x3s_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
    # backend.h mirror
    L.find_best_match.restype = C.c_size_t
    L.find_best_match.argtypes = [C.c_void_p]
    L.set_forward_window.argtypes = [C.c_size_t]
    L.get_forward_window.restype = C.c_size_t
    L.set_max_match_count.argtypes = [C.c_int]
    L.get_max_match_count.restype = C.c_int
    L.set_magic_factor1.argtypes = [C.c_size_t]
    L.get_magic_factor1.restype = C.c_size_t
    L.set_magic_factor2.argtypes = [C.c_size_t]
    L.get_magic_factor2.restype = C.c_size_t
    L.x3_search_prepare.argtypes = [C.c_void_p, C.c_size_t]
    L.x3_search_prepare.restype = None
    L.x3_search_release.restype = None
    L.x3_search_prepare_ms.restype = C.c_double
    L.x3_search_startup_ms.restype = C.c_double
    L.x3_search_landed_ms.restype = C.c_double
    L.x3_search_ready.restype = C.c_size_t
    L.x3_search_wait.restype = None
    L.x3s_search_host_stream.restype = C.c_int
    L.x3s_search_host_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.POINTER(Timing), C.POINTER(C.c_size_t)]
    L.x3_search_table.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.x3_backend_set_dict.argtypes = [C.c_void_p, C.c_void_p]
    _lib = L
    return L


def device_count() -> int:
    return int(lib().x3s_device_count())


def default_kernel(W: int = 8192, t: int = 15, want_table: bool = False) -> int:
    """The kernel KERNEL_DEFAULT resolves to for these parameters (x3s_default_kernel)."""
    return int(lib().x3s_default_kernel(W, t, int(want_table)))


def rank_profile(device: int = 0):
    """Per-kernel-family device time of the last rank search run with X3_RANK_PROFILE=1:
    {family: (ms, elements, launches)}."""
    out = {}
    for kind, name in enumerate(("radix", "level", "setup")):
        ms, el, nl = C.c_double(), C.c_double(), C.c_int()
        _check(lib().x3s_rank_profile(device, kind, C.byref(ms), C.byref(el), C.byref(nl)))
        out[name] = (ms.value, el.value, nl.value)
    return out


def rank_plan(n: int, W: int, lanes: int = 0):
    """(chunk_positions, chunks, lanes_used) of a rank search over n positions (x3s_rank_plan)."""
    ch, cnt, used = C.c_size_t(), C.c_size_t(), C.c_int()
    _check(lib().x3s_rank_plan(n, W, lanes, C.byref(ch), C.byref(cnt), C.byref(used)))
    return ch.value, cnt.value, used.value


def set_devices(ids):
    """Restricts search_host() to these CUDA ordinals (None/[] = all, in order)."""
    ids = list(ids or [])
    arr = (C.c_int * max(1, len(ids)))(*ids)
    _check(lib().x3s_set_devices(arr, len(ids)))


def required_bytes(n: int, W: int) -> int:
    return int(lib().x3s_required_bytes(n, W))


def _check(rc: int):
    if rc != X3S_OK:
        raise X3SearchError(rc, lib().x3s_last_error().decode(errors="replace"))


def padded(data, W: int) -> np.ndarray:
    """The reference's input buffer: the data followed by W zero bytes (x3.c:579,590)."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    out = np.zeros(len(a) + W, dtype=np.uint8)
    out[: len(a)] = a
    return out


def search_host(data, W: int = 8192, t: int = 15, ngpus: int = 1, variant: int = KERNEL_DEFAULT,
                want_table: bool = False, pinned: bool = False):
    """One call of the hot path over a host buffer through the C ABI.

    pinned=True stages the input and Lstar in page-locked memory (x3s_host_alloc), which is what lets
    the library pipeline a shard piece by piece (upload | search | copy back).

    Returns (lstar[n] u8, H[n,32] u8 or None, Timing)."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = len(a)
    H = np.empty((n, MAX_MATCH_LEN), dtype=np.uint8) if want_table else None
    tm = Timing()
    if pinned and n > 0:
        L = lib()
        hx, hl = L.x3s_host_alloc(n + W), L.x3s_host_alloc(n)
        if not hx or not hl:
            raise X3SearchError(X3S_ERR_CUDA, L.x3s_last_error().decode(errors="replace"))
        try:
            xv = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_uint8)), shape=(n + W,))
            xv[:n] = a
            xv[n:] = 0
            rc = L.x3s_search_host(hx, n, W, t, ngpus, variant, hl, H.ctypes.data if H is not None else None,
                                   C.byref(tm))
            _check(rc)
            lstar = np.ctypeslib.as_array(C.cast(hl, C.POINTER(C.c_uint8)), shape=(n,)).copy()
        finally:
            L.x3s_host_free(hx)
            L.x3s_host_free(hl)
        return lstar, H, tm
    x = padded(a, W)
    lstar = np.empty(n, dtype=np.uint8)
    rc = lib().x3s_search_host(x.ctypes.data, n, W, t, ngpus, variant, lstar.ctypes.data,
                               H.ctypes.data if H is not None else None, C.byref(tm))
    _check(rc)
    return lstar, H, tm


def search_device(device: int, d_x: int, n: int, W: int, t: int, d_lstar: int, d_H: int | None = None,
                  stream: int | None = None, variant: int = KERNEL_DEFAULT):
    """Asynchronous search over a device-resident buffer (raw device pointers)."""
    _check(lib().x3s_search_device(device, d_x, n, W, t, d_lstar, d_H, stream, variant))


class Backend:
    """Mirror of the reference's backend.h for one padded input buffer.

    Usage follows the reference's main(): setters first (x3.c:499-510), then
    prepare() where the reference has just called fload() (x3.c:591), then
    find_best_match(offset) from the parse loop (x3.c:383,400)."""

    def __init__(self):
        self.L = lib()
        self._buf = None

    def set_forward_window(self, n: int):
        self.L.set_forward_window(n)

    def get_forward_window(self) -> int:
        return int(self.L.get_forward_window())

    def set_max_match_count(self, n: int):
        self.L.set_max_match_count(n)

    def get_max_match_count(self) -> int:
        return int(self.L.get_max_match_count())

    def set_magic_factor1(self, f: int):
        self.L.set_magic_factor1(f)

    def get_magic_factor1(self) -> int:
        return int(self.L.get_magic_factor1())

    def set_magic_factor2(self, f: int):
        self.L.set_magic_factor2(f)

    def get_magic_factor2(self) -> int:
        return int(self.L.get_magic_factor2())

    def set_dict(self, find_ptr, len_ptr):
        """Registers dictionary queries (C function pointers or None)."""
        self.L.x3_backend_set_dict(find_ptr, len_ptr)

    def prepare(self, data) -> np.ndarray:
        self._buf = padded(data, self.get_forward_window())
        self._n = len(self._buf) - self.get_forward_window()
        self.L.x3_search_prepare(self._buf.ctypes.data, self._n)
        return self._buf

    def find_best_match(self, offset: int) -> int:
        return int(self.L.find_best_match(self._buf.ctypes.data + offset))

    def tables(self):
        H, Ls, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self.L.x3_search_table(C.byref(H), C.byref(Ls), C.byref(n))
        ls = np.ctypeslib.as_array(C.cast(Ls, C.POINTER(C.c_uint8)), shape=(n.value,)).copy() if Ls.value else None
        h = (np.ctypeslib.as_array(C.cast(H, C.POINTER(C.c_uint8)), shape=(n.value, MAX_MATCH_LEN)).copy()
             if H.value else None)
        return h, ls

    def release(self):
        self.L.x3_search_release()
        self._buf = None


def shard_ranges(n: int, world: int, align: int = 4096):
    """Contiguous position ranges [a_g, b_g) for `world` ranks.
    Rank g needs bytes [a_g, b_g + W - 2] of the padded buffer: a trailing halo."""
    cuts = [min(n, (n * g // world) // align * align) for g in range(world)] + [n]
    return [(cuts[g], cuts[g + 1]) for g in range(world)]
