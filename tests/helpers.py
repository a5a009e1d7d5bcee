"""Shared comparison helpers: GPU (C-ABI) output vs oracle output on the same records."""
from __future__ import annotations

import numpy as np

PRIMAL_TOL = 1e-6   # north_star: primal <= 1e-6 relative
KKT_TOL = 1e-6      # north_star: KKT residual <= 1e-6


def rel_inf(a, b):
    """||a-b||_inf / max(1, ||b||_inf), row-wise (SURVEY.md 8(c) primal parity)."""
    return np.abs(a - b).max(axis=1) / np.maximum(1.0, np.abs(b).max(axis=1))


def strongly_active(y, scale_tol=1e-8):
    """Rows whose multiplier is significant: |y| > 1e-8 * max(1, ||y||_inf)."""
    sc = np.maximum(1.0, np.abs(y).max(axis=1, keepdims=True))
    return np.abs(y) > scale_tol * sc


def compare(L, gpu: dict, ora: dict, gdiag: dict | None = None, odiag: dict | None = None) -> dict:
    ok = (gpu["status"] == 0) & (ora["status"] == 0)
    r = dict(n=len(ok), both_ok=int(ok.sum()), status_equal=bool((gpu["status"] == ora["status"]).all()))
    if ok.any():
        r["primal"] = float(rel_inf(gpu["x"][ok], ora["x"][ok]).max())
        r["tau"] = float(rel_inf(gpu["tau"][ok], ora["tau"][ok]).max())
        r["kkt_gpu"] = float(gpu["kkt"][ok].max())
        r["kkt_oracle"] = float(ora["kkt"][ok].max())
        r["mask_equal"] = float((gpu["active"][ok] == ora["active"][ok]).all(axis=1).mean())
        if gdiag is not None and odiag is not None:
            r["eopt"] = float(rel_inf(gdiag["eopt"][ok], odiag["eopt"][ok]).max())
            r["y1"] = float(rel_inf(gdiag["y1"][ok], odiag["y1"][ok]).max())
            sa_g, sa_o = strongly_active(gdiag["y1"][ok]), strongly_active(odiag["y1"][ok])
            r["strong_active_equal"] = float((sa_g == sa_o).all(axis=1).mean())
            sign_ok = np.sign(gdiag["y1"][ok]) * sa_o == np.sign(odiag["y1"][ok]) * sa_o
            r["strong_sign_equal"] = float(sign_ok.all(axis=1).mean())
    return r


def mask_bits(mask_words, n_rows):
    """(B,4) uint32 -> (B,n_rows) bool."""
    bits = np.unpackbits(np.ascontiguousarray(mask_words).view(np.uint8), axis=1, bitorder="little")
    return bits[:, :n_rows].astype(bool)


def mask_differences_are_degenerate(desc, L, recs, x, x0, mask_a, mask_b, rel_tol=1e-5):
    """Active sets may differ only on rows that are TIGHT at the solution (weakly active / implied rows at a
    degenerate vertex, where the multipliers are not unique).  Returns (n_differing_problems, all_tight)."""
    from tests.assemble_np import level_matrices
    a, b = mask_bits(mask_a, L.n_rows), mask_bits(mask_b, L.n_rows)
    diff = np.nonzero((a != b).any(axis=1))[0]
    ok = True
    for i in diff:
        _, _, C, lA, uA, _ = level_matrices(desc, recs[i], 1, x0[i])
        cx = C @ x[i]
        rows = np.nonzero(a[i] != b[i])[0]
        slack = np.minimum(np.abs(cx[rows] - lA[rows]), np.abs(uA[rows] - cx[rows]))
        ok &= bool((slack <= rel_tol * np.maximum(1.0, np.abs(cx[rows]))).all())
    return len(diff), ok
