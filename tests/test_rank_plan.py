"""CPU tests of how the rank search cuts an input into chunks and lanes (x3s_rank_plan: host logic of
x3_search_rank.cu, no device needed)."""
import os

import pytest

CHUNK_LIMIT = 1 << 24  # chunk + distances must stay below: ranks and positions are 24-bit fields


@pytest.fixture(autouse=True)
def _no_knobs(monkeypatch):
    monkeypatch.delenv("X3_RANK_LANES", raising=False)
    monkeypatch.delenv("X3_RANK_PROFILE", raising=False)


@pytest.mark.parametrize("n", [1, 4095, 4096, 65536, 10_192_446, 16_769_024, 16_769_025, 17_400_000, 50_000_000,
                               211_938_580, 1 << 32])
@pytest.mark.parametrize("W", [0, 33, 34, 8192, 65536, 1 << 20, (1 << 23) + 33])
@pytest.mark.parametrize("lanes", [0, 1, 2, 3, 4, 8, 50])
def test_chunks_cover_the_input_and_fit_24_bits(pkg, n, W, lanes):
    chunk, chunks, used = pkg.rank_plan(n, W, lanes)
    D = max(W - 33, 0)
    assert chunk % 4096 == 0 and chunk > 0
    assert chunk + D < CHUNK_LIMIT
    assert (chunks - 1) * chunk < n <= chunks * chunk          # the chunks tile [0, n) exactly, none is empty
    assert 1 <= used <= min(8, chunks)
    if lanes >= 1:
        assert used <= lanes
    # as few chunks as the element format allows, unless more lanes were asked for
    cmax = (CHUNK_LIMIT - 1 - D) & ~4095
    least = -(-n // cmax)
    assert chunks >= least
    if lanes in (0, 1):
        assert chunks == least


def test_default_is_one_lane_per_chunk_up_to_four(pkg):
    assert pkg.rank_plan(10_192_446, 8192)[1:] == (1, 1)        # C2, C4: one chunk, one lane
    assert pkg.rank_plan(17_400_000, 8192)[1:] == (2, 2)
    assert pkg.rank_plan(50_000_000, 1 << 20)[1:] == (4, 4)     # C3: each chunk carries a 1 MB halo
    assert pkg.rank_plan(211_938_580, 8192)[1:] == (13, 4)      # C5


def test_knob_and_profile_mode(pkg, monkeypatch):
    monkeypatch.setenv("X3_RANK_LANES", "3")
    assert pkg.rank_plan(10_192_446, 8192)[1:] == (3, 3)
    assert pkg.rank_plan(100_000, 8192)[1:] == (1, 1)           # tiny inputs are not split
    monkeypatch.setenv("X3_RANK_PROFILE", "1")                  # events bracket the launches of one stream
    assert pkg.rank_plan(211_938_580, 8192)[2] == 1


def test_windows_beyond_the_rank_search_are_refused(pkg):
    with pytest.raises(pkg.X3SearchError) as ei:
        pkg.rank_plan(1000, (1 << 23) + 34)
    assert ei.value.code == pkg.X3S_ERR_UNSUPP
