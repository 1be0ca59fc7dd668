"""Shared test helpers (tie-/discontinuity-aware comparison of decoded detections)."""
import numpy as np


def match_rows(got, ref, tol=1e-3, box_tol=None):
    """got/ref: (K, C) detection rows of ONE image.  Rows are matched greedily on the bbox
    (first 4 columns) + score; returns (fraction of ref rows matched,
    fraction of elements within tol among matched rows)."""
    box_tol = tol if box_tol is None else box_tol
    used = np.zeros(len(got), bool)
    matched = 0; ok = 0; tot = 0
    for r in ref:
        d = np.abs(got[:, :5] - r[None, :5]).max(axis=1)
        d[used] = np.inf
        j = int(np.argmin(d))
        if d[j] <= box_tol:
            used[j] = True; matched += 1
            e = np.abs(got[j] - r) <= tol * np.maximum(1.0, np.abs(r))
            ok += int(e.sum()); tot += e.size
    return matched / max(1, len(ref)), (ok / tot if tot else 0.0)
