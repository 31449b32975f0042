"""ctypes access to the oracle (oracle/build/libx3oracle.so) and to the compiled,
unmodified reference (oracle/_ref/libx3ref.so).  TEST INFRASTRUCTURE: imported only
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "build" / "libx3oracle.so"
REF_DIR = ROOT / "oracle" / "_ref"
REF_SO = REF_DIR / "libx3ref.so"

DICT_FIND_FN = C.CFUNCTYPE(C.c_size_t, C.c_void_p)
DICT_LEN_FN = C.CFUNCTYPE(C.c_size_t, C.c_size_t)
SIZE32 = C.c_size_t * 32

_ora = None
_ref = None


def oracle() -> C.CDLL:
    global _ora
    if _ora is None:
        if not ORACLE_SO.exists():
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            subprocess.run(["make", "-C", str(ROOT / "oracle"), "oracle"], check=True, env=env,
                           stdout=subprocess.DEVNULL)
        L = C.CDLL(str(ORACLE_SO))
        L.x3o_histogram.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.x3o_select.restype = C.c_size_t
        L.x3o_select.argtypes = [C.POINTER(C.c_size_t), C.c_void_p, C.c_int, C.c_size_t, C.c_size_t,
                                 C.c_void_p, C.c_void_p]
        L.x3o_find_best_match.restype = C.c_size_t
        L.x3o_find_best_match.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t,
                                          C.c_void_p, C.c_void_p]
        L.x3o_lstar_from_count.restype = C.c_uint8
        L.x3o_lstar_from_count.argtypes = [C.POINTER(C.c_size_t), C.c_int]
        L.x3o_filter_from_lstar.restype = C.c_size_t
        L.x3o_filter_from_lstar.argtypes = [C.c_uint8, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_void_p]
        L.x3o_table_plain.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.x3o_table_fast.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.x3o_call_range.restype = C.c_uint64
        L.x3o_call_range.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
        _ora = L
    return _ora


def have_ref() -> bool:
    return REF_SO.exists()


def ref() -> C.CDLL:
    """The compiled reference backend.c + dict.c (reference backend.h / dict.h symbols)."""
    global _ref
    if _ref is None:
        L = C.CDLL(str(REF_SO))
        L.find_best_match.restype = C.c_size_t
        L.find_best_match.argtypes = [C.c_void_p]
        L.set_forward_window.argtypes = [C.c_size_t]
        L.set_max_match_count.argtypes = [C.c_int]
        L.set_magic_factor1.argtypes = [C.c_size_t]
        L.set_magic_factor2.argtypes = [C.c_size_t]
        L.dict_find_match.restype = C.c_size_t
        L.dict_find_match.argtypes = [C.c_void_p]
        L.dict_get_len_by_index.restype = C.c_size_t
        L.dict_get_len_by_index.argtypes = [C.c_size_t]
        L.dict_get_elems.restype = C.c_size_t
        L.dict_can_insert_elem.restype = C.c_int
        L.dict_insert_elem.argtypes = [C.c_void_p]
        L.dict_update_costs.argtypes = [C.c_void_p]
        L.dict_query_elem.restype = C.c_int
        L.dict_query_elem.argtypes = [C.c_void_p]
        L.elem_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _ref = L
    return _ref


class RefElem(C.Structure):
    """reference dict.h:7-13"""
    _fields_ = [("s", C.c_char * 32), ("len", C.c_size_t), ("last_pos", C.c_void_p), ("cost", C.c_size_t),
                ("tag", C.c_size_t)]


def ref_dict_insert(buf: np.ndarray, offset: int, length: int) -> bool:
    """Inserts buf[offset:offset+length] into the compiled reference's dictionary the
    way compress() does (reference x3.c:408-418)."""
    R = ref()
    e = RefElem()
    R.elem_fill(C.byref(e), buf.ctypes.data + offset, length)
    if R.dict_query_elem(C.byref(e)) != 0:
        return False
    if not R.dict_can_insert_elem():
        R.dict_enlarge()
    R.dict_insert_elem(C.byref(e))
    return True


def padded(data, W: int) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    out = np.zeros(len(a) + W + 64, dtype=np.uint8)
    out[: len(a)] = a
    return out


def table(data, W: int, t: int, p0: int = 0, p1: int | None = None, plain: bool = False, h16: bool = False):
    """(H[n,32] u8, Lstar[n] u8) for positions [p0, p1) of data padded with W zeros."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = len(a)
    p1 = n if p1 is None else p1
    x = padded(a, W)
    m = p1 - p0
    H = np.zeros((m, 32), dtype=np.uint16 if h16 else np.uint8)
    ls = np.zeros(m, dtype=np.uint8)
    h8 = None if h16 else H.ctypes.data
    h16p = H.ctypes.data if h16 else None
    if m > 0:
        if plain:
            oracle().x3o_table_plain(x.ctypes.data, p0, p1, W, t, h8, h16p, ls.ctypes.data)
        else:
            oracle().x3o_table_fast(x.ctypes.data, len(x), p0, p1, W, t, h8, h16p, ls.ctypes.data)
    return H, ls


def histogram(x: np.ndarray, p: int, W: int) -> np.ndarray:
    cnt = SIZE32()
    oracle().x3o_histogram(x.ctypes.data + p, W, cnt)
    return np.array(list(cnt), dtype=np.uint64)


def lstar_from_count(count, t: int) -> int:
    cnt = SIZE32(*[int(v) for v in count])
    return int(oracle().x3o_lstar_from_count(cnt, t))
