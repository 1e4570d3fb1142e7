"""Comparison of device sync positions with the oracle's, aware of the one legitimate way they can differ.

The reference picker (decode.rs:236-262) keeps a candidate only if its correlation is STRICTLY larger than the current
peak's.  The device correlation is the same sum in a different association (and `f` went through FMAs), so it differs from
the reference's sequential f32 sum by ~1e-7 relative.  Where two neighbouring candidates are closer than that -- 1 ulp apart
in the reference's own arithmetic -- which of them wins is decided by the last bit, on the CPU as on the GPU: a TIE.  A tied
position may differ by one work sample; everything else must be identical.
"""
import numpy as np

import oracle

TIE_MARGIN = 1e-6      # relative difference of the two correlation values below which the reference's own choice is noise


def sync_ties(pos_gpu, pos_ref, filtered, work_rate):
    """Asserts that the positions are the oracle's up to ties; returns the indices of the tied rows (usually empty)."""
    pos_gpu = np.asarray(pos_gpu).astype(np.int64)
    pos_ref = np.asarray(pos_ref).astype(np.int64)
    assert pos_gpu.size == pos_ref.size, (pos_gpu.size, pos_ref.size)
    bad = np.nonzero(pos_gpu != pos_ref)[0]
    if bad.size == 0:
        return []
    _, corr = oracle.find_sync(filtered, work_rate, want_corr=True)
    for j in bad:
        a, b = int(pos_ref[j]), int(pos_gpu[j])
        assert abs(a - b) <= 1, f"sync position {j}: {b} vs oracle {a}"
        margin = abs(float(corr[a]) - float(corr[b])) / max(abs(float(corr[a])), 1e-30)
        assert margin <= TIE_MARGIN, f"sync position {j}: {b} vs oracle {a}, correlation margin {margin:.2e} is not a tie"
    return [int(j) for j in bad]


def rows_without(rows, ties, px=2080):
    keep = np.ones(rows.size // px, dtype=bool)
    for j in ties:
        if j < keep.size:
            keep[j] = False
    return rows.reshape(-1, px)[keep]
