"""Index algebra of the round-2 kernels, emulated in numpy against the plain definition (no GPU needed).

These mirror the loops of kernels_sync2.cuh line by line (window indices, tap-pair tables, alignment variants, the static
tile/pool bookkeeping), so that a change of a constant there has a CPU test to break first; the kernels themselves are
checked against the oracle by the -m gpu tests.  Also: the tie-aware comparison helper of tests/_parity.py on crafted data.
"""
import numpy as np
import pytest

import oracle
from _parity import rows_without, sync_ties


def lp_tables(c, dec=0):
    """launch.cu: make_lp_taps -- p[j+1] = (c[j], c[j+1]), pd[j+dec] = (c[j], c[j+dec]); taps outside [0, NT) are zero."""
    nt = len(c)
    tap = lambda j: float(c[j]) if 0 <= j < nt else 0.0
    p = [(tap(j), tap(j + 1)) for j in range(-1, 63)]
    pd = [(tap(j), tap(j + dec)) for j in range(-dec, 72 - dec)]
    return p, pd


@pytest.mark.parametrize("nt", [37, 43, 61])
def test_lowpass_phase1_pair_algebra(nt):
    """k_lowpass_records phase 1: one window sample x the taps of two neighbouring outputs."""
    rng = np.random.default_rng(nt)
    c = rng.standard_normal(nt)
    eoff = (nt - 1 + 3) // 4 * 4
    wn = eoff + 32
    w = rng.standard_normal(wn)
    p, _ = lp_tables(c)
    fr = np.zeros(32)
    for h in range(2):
        acc = np.zeros((8, 2))
        for j in range(-1, nt):
            t = p[j + 1]
            for v in range(8):
                m = eoff + 2 * (8 * h + v) - j
                assert 0 <= m < wn
                acc[v, 0] += w[m] * t[0]
                acc[v, 1] += w[m] * t[1]
        for v in range(8):
            fr[16 * h + 2 * v], fr[16 * h + 2 * v + 1] = acc[v]
    want = np.array([sum(c[jj] * w[eoff + o - jj] for jj in range(nt)) for o in range(32)])
    assert np.allclose(fr, want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("nt,dec", [(37, 3), (43, 4), (61, 5)])
@pytest.mark.parametrize("off", [0, 1, 2, 3])
def test_gather_quad_alignment_variants(nt, dec, off):
    """k_gather_rows_lp / gather_quad<NT, DEC, OFF>: four pixels from a 16-byte aligned window that starts OFF samples early."""
    rng = np.random.default_rng(100 * nt + off)
    c = rng.standard_normal(nt)
    eoff = (nt - 1 + 3) // 4 * 4
    win = (off + eoff + 3 * dec + 1 + 3) // 4 * 4
    assert eoff - (nt - 1) >= 0 and off + eoff + 3 * dec < win and nt + dec <= 72
    w = rng.standard_normal(win)
    _, pd = lp_tables(c, dec)
    acc2 = np.zeros((2, 2))
    for j in range(-dec, nt):
        t = pd[j + dec]
        for pp in range(2):
            m = off + eoff + dec * (2 * pp) - j
            assert 0 <= m < win
            acc2[pp, 0] += w[m] * t[0]
            acc2[pp, 1] += w[m] * t[1]
    got = acc2.reshape(4)
    want = np.array([sum(c[jj] * w[off + eoff + dec * u - jj] for jj in range(nt)) for u in range(4)])
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("dec,px", [(3, 2080), (4, 2080), (5, 2080)])
def test_gather_span_covers_every_thread_window(dec, px):
    """The staged span of a half row (launch_gather_lp / the kernel's `span`) holds the window of its last thread for every
    alignment, and the aligned origin reproduces the row's samples."""
    nt = {3: 37, 4: 43, 5: 61}[dec]
    eoff = (nt - 1 + 3) // 4 * 4
    part_px = (px // 2 + 3) // 4 * 4
    span = dec * part_px + eoff + 12
    assert span % 4 == 0
    for pos in range(0, 64):                       # any row position
        for part in range(2):
            g0 = pos + dec * (part * part_px) - eoff
            a0 = g0 & ~3                            # floor to a multiple of 4, also for negative g0 (Python ints behave like the kernel's long long)
            off = g0 - a0
            assert 0 <= off <= 3 and a0 % 4 == 0
            c4_last = part_px - 4
            win = (off + eoff + 3 * dec + 1 + 3) // 4 * 4
            assert dec * c4_last + win <= span
            # pixel c of the part, tap jj reads sample pos + dec*(part*part_px + c) - jj = a0 + (dec*c + off + eoff - jj)
            c, jj = 7, 5
            assert a0 + (dec * c + off + eoff - jj) == pos + dec * (part * part_px + c) - jj


def test_record_pool_regions_are_disjoint():
    """k_lowpass_records: a tile stores into its own region unless it has more records than the region holds; then into the
    overflow area behind the regions (api.cu: pool_cap >= max_tiles * region + overflow)."""
    region, ntiles, max_tiles = 640, 5850, 5851
    overflow = max(11_231_877 // 8, 1 << 16)
    pool_cap = max_tiles * region + overflow
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 400, ntiles)
    counts[[3, 77, 5000]] = [900, 3840, 641]       # three tiles that do not fit their region
    cursor, used = 0, []
    for t, n in enumerate(counts):
        if n > region:
            base = ntiles * region + cursor
            cursor += n
        else:
            base = t * region
        assert base + n <= pool_cap
        used.append((base, base + n))
    used.sort()
    for (a0, a1), (b0, b1) in zip(used, used[1:]):
        assert a1 <= b0


def test_static_tile_dealing_covers_every_tile_once():
    """k_lowpass_records: CTA b takes the tile groups b, b + grid, ...; a group is NW consecutive tiles, one per warp."""
    for ntiles, grid, nw in [(5850, 2960, 1), (5850, 366, 8), (260, 260, 1), (7, 3, 4)]:
        seen = np.zeros(ntiles, int)
        for b in range(grid):
            it = 0
            while (b + it * grid) * nw < ntiles:   # the loop condition: the group's first tile exists
                for warp in range(nw):
                    tile = (b + it * grid) * nw + warp
                    if tile < ntiles:               # `active`
                        seen[tile] += 1
                it += 1
        assert np.all(seen == 1)


def test_sync_ties_helper_accepts_only_ties():
    """tests/_parity.py on crafted data: equal -> no ties; one sample off with equal correlation -> tie; anything else fails."""
    work_rate = 12480
    rng = np.random.default_rng(3)
    f = rng.standard_normal(20 * 6240).astype(np.float32) * 0.01 + 1.0
    pos_ref, corr = oracle.find_sync(f, work_rate, want_corr=True)
    assert pos_ref.size >= 3
    assert sync_ties(pos_ref, pos_ref, f, work_rate) == []
    moved = pos_ref.astype(np.int64).copy()
    moved[2] += 1
    a, b = int(pos_ref[2]), int(moved[2])
    margin = abs(float(corr[a]) - float(corr[b])) / abs(float(corr[a]))
    if margin <= 1e-6:
        assert sync_ties(moved, pos_ref, f, work_rate) == [2]
    else:
        with pytest.raises(AssertionError):
            sync_ties(moved, pos_ref, f, work_rate)
    far = pos_ref.astype(np.int64).copy()
    far[1] += 5
    with pytest.raises(AssertionError):
        sync_ties(far, pos_ref, f, work_rate)
    # a constant signal: every correlation value is the same, so ANY neighbour is a tie of the strict `>`
    flat = np.ones(8 * 6240, dtype=np.float32)
    pos_flat, corr_flat = oracle.find_sync(flat, work_rate, want_corr=True)
    assert np.all(corr_flat == corr_flat[0]) and pos_flat.size >= 3
    shifted = pos_flat.astype(np.int64).copy()
    shifted[1] += 1
    assert sync_ties(shifted, pos_flat, flat, work_rate) == [1]
    rows = np.arange(4 * 2080, dtype=np.float32)
    assert rows_without(rows, [1]).shape == (3, 2080)
