"""Independent QP checkers used to pin the oracle (no code shared with oracle/ or qppvm_b200/):
HiGHS 1.12 QP (scipy-bundled), brute-force active-set enumeration, numpy KKT certificate."""
from __future__ import annotations

import itertools

import numpy as np
import scipy.sparse as sp

INF = 1e20


def highs_qp(Q, c, C, lA, uA):
    """min 1/2 x'Qx + c'x s.t. lA <= Cx <= uA.  Returns (optimal?, x, row_dual)."""
    from scipy.optimize._highspy import _core as hc
    n, m = len(c), C.shape[0]
    inf = hc.kHighsInf
    model = hc.HighsModel()
    lp = model.lp_
    lp.num_col_, lp.num_row_ = n, m
    lp.col_cost_ = np.asarray(c, float)
    lp.col_lower_ = np.full(n, -inf)
    lp.col_upper_ = np.full(n, inf)
    lp.row_lower_ = np.where(lA <= -0.5 * INF, -inf, lA)
    lp.row_upper_ = np.where(uA >= 0.5 * INF, inf, uA)
    Cs = sp.csc_matrix(C)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.num_col_, lp.a_matrix_.num_row_ = n, m
    lp.a_matrix_.start_ = Cs.indptr.astype(np.int32)
    lp.a_matrix_.index_ = Cs.indices.astype(np.int32)
    lp.a_matrix_.value_ = Cs.data
    Ql = sp.csc_matrix(np.tril(Q))
    h = model.hessian_
    h.dim_, h.format_ = n, hc.HessianFormat.kTriangular
    h.start_ = Ql.indptr.astype(np.int32)
    h.index_ = Ql.indices.astype(np.int32)
    h.value_ = Ql.data
    H = hc._Highs()
    H.setOptionValue("output_flag", False)
    H.setOptionValue("primal_feasibility_tolerance", 1e-10)
    H.setOptionValue("dual_feasibility_tolerance", 1e-10)
    H.setOptionValue("time_limit", 5.0)      # HiGHS can cycle on the cond~1e9 levels
    H.passModel(model)
    H.run()
    sol = H.getSolution()
    return H.getModelStatus() == hc.HighsModelStatus.kOptimal, np.array(sol.col_value), np.array(sol.row_dual)


def kkt_numpy(H, g, C, lA, uA, x, y):
    """SURVEY.md 8(c) residual in plain numpy: (r_stat, r_prim, r_comp) scaled."""
    Hx = H @ x
    cx = C @ x if len(lA) else np.zeros(0)
    r_stat = np.abs(Hx + g - (C.T @ y if len(lA) else 0)).max() / max(1.0, np.abs(g).max(), np.abs(Hx).max())
    viol = np.maximum(0.0, np.maximum(lA - cx, cx - uA)).max() if len(lA) else 0.0
    r_prim = viol / max(1.0, np.abs(x).max(), np.abs(cx).max() if len(lA) else 0.0)
    ineq = lA != uA
    comp = np.where(y > 0, y * np.abs(cx - lA), np.where(y < 0, -y * np.abs(uA - cx), 0.0))
    r_comp = (comp[ineq].max() if ineq.any() else 0.0) / (max(1.0, np.abs(y).max() if len(y) else 0) * max(1.0, np.abs(cx).max() if len(lA) else 0))
    return r_stat, r_prim, r_comp


def brute_force_qp(H, g, C, lA, uA, tol=1e-9):
    """Enumerates every working set (each inequality row inactive / at lA / at uA; equalities always active)
    and returns the KKT point (x, y).  H must be positive definite.  Exponential: toy sizes only."""
    n, m = len(g), len(lA)
    eq = [i for i in range(m) if lA[i] == uA[i]]
    iq = [i for i in range(m) if lA[i] != uA[i]]
    best = None
    for states in itertools.product((0, 1, 2), repeat=len(iq)):
        rows = list(eq) + [i for i, s in zip(iq, states) if s]
        rhs = [lA[i] for i in eq] + [lA[i] if s == 1 else uA[i] for i, s in zip(iq, states) if s]
        if any(abs(v) >= 0.5 * INF for v in rhs) or len(rows) > n:
            continue
        k = len(rows)
        K = np.zeros((n + k, n + k))
        K[:n, :n] = H
        K[:n, n:] = -C[rows].T
        K[n:, :n] = C[rows]
        try:
            sol = np.linalg.solve(K, np.concatenate([-g, rhs]))
        except np.linalg.LinAlgError:
            continue
        x, lam = sol[:n], sol[n:]
        cx = C @ x
        if (cx < lA - tol).any() or (cx > uA + tol).any():
            continue
        good = True
        for j, (i, s) in enumerate([(i, s) for i, s in zip(iq, states) if s]):
            l = lam[len(eq) + j]
            if (s == 1 and l < -tol) or (s == 2 and l > tol):
                good = False
        if not good:
            continue
        y = np.zeros(m)
        y[rows] = lam
        f = 0.5 * x @ H @ x + g @ x
        if best is None or f < best[2] - 1e-12:
            best = (x, y, f)
    return best
