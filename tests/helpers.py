"""Shared checks for the parity tests (tests/ only)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Pi values: BASELINE.json north_star asks for 1e-5 relative; fp64 paths that differ only in
# summation order agree far tighter, so the tests hold them to 1e-9 and report the worst seen.
VAL_RTOL = 1e-9
# tie band for top-k membership: k-th and (k+1)-th values closer than this (relative) may swap
TIE_RTOL = 1e-9


def load_graph(name):
    z = np.load(os.path.join(GOLDEN, f"graph_{name}.npz"))
    return z["indptr"], z["indices"]


def check_topk_rows(indptr, indices, node_idx, coef, rmax, K, col, val, row=None,
                    val_rtol=VAL_RTOL, tie_rtol=TIE_RTOL, max_rows=None):
    """Check [S*K] slot arrays against the oracle's full reserve rows.

    Per source: (1) the number of filled slots is min(K, #positive reserve entries);
    (2) every (col, val) matches the oracle reserve value within val_rtol; (3) no duplicates;
    (4) membership: every selected value >= t*(1-tie) and every reserve value > t*(1+tie) is
    selected, t = k-th largest oracle value -- i.e. index sets are identical wherever the
    k-th/(k+1)-th values differ by more than the band (north_star); (5) row_idx == source."""
    from oracle import gfpush as og
    col = np.asarray(col).reshape(-1, K)
    val = np.asarray(val).reshape(-1, K)
    S = len(node_idx)
    assert col.shape[0] == S
    worst = 0.0
    rows = range(S) if max_rows is None else np.linspace(0, S - 1, min(S, max_rows)).astype(int)
    for it in rows:
        src = int(node_idx[it])
        dense, _, _ = og.reserve_row(indptr, indices, src, coef, rmax)
        filled = val[it] > 0
        c, v = col[it][filled], val[it][filled]
        npos = int((dense > 0).sum())
        k = min(K, npos)
        assert filled.sum() == k, f"row {it} (src {src}): {filled.sum()} filled, expected {k}"
        assert len(np.unique(c)) == len(c), f"row {it}: duplicate columns"
        ref = dense[c]
        assert np.all(ref > 0), f"row {it}: selected a node with zero reserve"
        rel = np.abs(v - ref) / ref
        worst = max(worst, float(rel.max()) if len(rel) else 0.0)
        assert np.all(rel <= val_rtol), f"row {it}: value mismatch {rel.max():.3e}"
        if k > 0 and npos > k:
            t = np.partition(dense, -k)[-k]
            assert np.all(ref >= t * (1 - tie_rtol)), f"row {it}: selected below the k-th value"
            must = np.nonzero(dense > t * (1 + tie_rtol))[0]
            assert np.isin(must, c).all(), f"row {it}: missed an entry above the k-th value"
        if row is not None:
            r = np.asarray(row).reshape(-1, K)[it]
            assert np.all(r[filled] == src)
            assert np.all(r[~filled] == 0) and np.all(col[it][~filled] == 0)
    return worst
