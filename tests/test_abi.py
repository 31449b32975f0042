"""CPU tests of the boundary: the library loads, exports every symbol the headers
declare, mirrors the reference's defaults, and refuses to compute without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: Path):
    txt = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    names = []
    for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", txt, flags=re.M):
        if "typedef" not in m.group(0):
            names.append(m.group(1))
    return names


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    declared = _declared(ROOT / "include" / "x3_search.h") + _declared(ROOT / "include" / "x3_backend.h")
    assert {"find_best_match", "set_forward_window", "get_forward_window", "set_max_match_count",
            "get_max_match_count", "get_magic_factor1", "set_magic_factor1", "get_magic_factor2",
            "set_magic_factor2", "x3_search_prepare", "x3_search_release", "x3_search_table",
            "x3s_search_host", "x3s_search_device"} <= set(declared)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/*.h but not exported"


def test_backend_defaults_match_reference(pkg):
    """reference backend.c:8,21,33,34"""
    b = pkg.Backend()
    assert b.get_forward_window() == 8192
    assert b.get_max_match_count() == 15
    assert b.get_magic_factor1() == 4
    assert b.get_magic_factor2() == 0
    b.set_forward_window(4096); b.set_max_match_count(3); b.set_magic_factor1(0); b.set_magic_factor2(2)
    assert (b.get_forward_window(), b.get_max_match_count(), b.get_magic_factor1(), b.get_magic_factor2()) == \
        (4096, 3, 0, 2)
    b.set_forward_window(8192); b.set_max_match_count(15); b.set_magic_factor1(4); b.set_magic_factor2(0)


def test_required_bytes_covers_padding(pkg):
    for n, W in [(0, 0), (1, 8192), (10_192_446, 8192), (5000, 1 << 20)]:
        assert pkg.required_bytes(n, W) >= n + W


def test_no_cpu_fallback(pkg):
    """Without a device the compute entry points must fail loudly, not compute."""
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    data = np.zeros(1000, dtype=np.uint8)
    with pytest.raises(pkg.X3SearchError) as ei:
        pkg.search_host(data, W=8192, t=15)
    assert ei.value.code == pkg.X3S_ERR_CUDA


def test_unsupported_parameters_are_rejected(pkg):
    """t >= 255 is refused only where u8 cells count (the 32-bin table, the brute-force variants);
    the Lstar-only search takes any t, like the reference (backend.c:21-26)."""
    data = np.zeros(1000, dtype=np.uint8)
    with pytest.raises(pkg.X3SearchError) as ei:
        pkg.search_host(data, W=8192, t=255, want_table=True)
    assert ei.value.code == pkg.X3S_ERR_UNSUPP
    with pytest.raises(pkg.X3SearchError) as ei:
        pkg.search_host(data, W=8192, t=300, variant=pkg.KERNEL_STREAM)
    assert ei.value.code == pkg.X3S_ERR_UNSUPP
    if pkg.device_count() == 0:
        with pytest.raises(pkg.X3SearchError) as ei:
            pkg.search_host(data, W=8192, t=300)  # accepted as a parameter; fails only for want of a device
        assert ei.value.code == pkg.X3S_ERR_CUDA


def test_shard_ranges_cover_input(pkg):
    for n in (0, 1, 4095, 4096, 10_192_446, 211_938_580):
        for world in (1, 2, 3, 4, 8):
            r = pkg.shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a <= b for a, b in r)
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert all(a % 16 == 0 for a, _ in r)
