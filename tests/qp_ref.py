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


def primal_active_set(A, b, C, lA, uA, eps, xp=None, maxit=4000):
    """min 1/2 ||A x - b||^2 + eps/2 ||x - xp||^2  s.t. lA <= C x <= uA  by a PRIMAL active-set method (the family qpOASES
    belongs to; the oracle and the CUDA kernels are DUAL methods, Goldfarb-Idnani): start from a feasible vertex found
    by HiGHS' LP phase 1, keep every iterate feasible, solve the equality-constrained problem on the working set, step
    until a constraint blocks (it joins the working set) or the step is complete (then the most wrong-signed multiplier
    leaves).  Whitened coordinates u = R x with R from a QR of [A; sqrt(eps) I] so that H is never formed; the working
    set is re-factorised from scratch every iteration (numpy QR): nothing incremental, nothing shared with oracle/.
    Returns (x, y, iterations); y > 0: active at lA, y < 0: active at uA (qpOASES sign convention)."""
    from scipy.linalg import solve_triangular
    from scipy.optimize import linprog
    m, n = C.shape
    Rq = np.linalg.qr(np.vstack([A, np.sqrt(eps) * np.eye(n)]), mode="r")
    u0 = solve_triangular(Rq.T, A.T @ b + (eps * xp if xp is not None else 0.0), lower=True)
    Wm = solve_triangular(Rq.T, C.T, lower=True)               # whitened normals, one column per row of C
    wnorm = np.linalg.norm(Wm, axis=0)
    eq = lA == uA
    lo_f, hi_f = (lA > -0.5 * INF) & ~eq, (uA < 0.5 * INF) & ~eq
    lp = linprog(np.zeros(n), A_ub=np.vstack([C[hi_f], -C[lo_f]]), b_ub=np.concatenate([uA[hi_f], -lA[lo_f]]),
                 A_eq=C[eq] if eq.any() else None, b_eq=lA[eq] if eq.any() else None, bounds=(None, None), method="highs")
    if lp.status != 0:
        raise RuntimeError("phase 1: " + lp.message)
    u = Rq @ lp.x
    work = [(int(i), 0) for i in np.nonzero(eq)[0]]
    for it in range(maxit):
        idx = [i for i, _ in work]
        if idx:
            Wk = Wm[:, idx]
            bk = np.array([uA[i] if s < 0 else lA[i] for i, s in work])
            Q, R1 = np.linalg.qr(Wk)
            ustar = u0 + Q @ solve_triangular(R1.T, bk - Wk.T @ u0, lower=True)
        else:
            ustar = u0
        p = ustar - u
        if len(idx) >= n or np.linalg.norm(p) <= 1e-11 * max(1.0, np.linalg.norm(u)):   # (n independent rows: a vertex)
            u = ustar
            y = solve_triangular(R1, Q.T @ (ustar - u0)) if idx else np.zeros(0)
            sc = max(1.0, np.abs(y).max()) if idx else 1.0
            worst, wj = 0.0, -1
            for j, (i, s) in enumerate(work):
                if s != 0 and y[j] * s < -1e-11 * sc and y[j] * s < worst:
                    worst, wj = y[j] * s, j
            if wj < 0:
                yy = np.zeros(m)
                yy[idx] = y
                return solve_triangular(Rq, ustar), yy, it
            del work[wj]
            continue
        cu, cp = Wm.T @ u, Wm.T @ p
        inw = np.zeros(m, dtype=bool)
        inw[idx] = True
        alpha, blk = 1.0, None
        thr = 1e-13 * np.linalg.norm(p) * wnorm                # a row only blocks if the step really moves along it
        for i in np.nonzero(~inw & ~eq)[0]:
            if lo_f[i] and cp[i] < -thr[i]:
                a = max((lA[i] - cu[i]) / cp[i], 0.0)
                if a < alpha:
                    alpha, blk = a, (int(i), 1)
            if hi_f[i] and cp[i] > thr[i]:
                a = max((uA[i] - cu[i]) / cp[i], 0.0)
                if a < alpha:
                    alpha, blk = a, (int(i), -1)
        u = u + alpha * p
        if blk is not None:
            work.append(blk)
    raise RuntimeError("primal active set: iteration limit")


def cascade_primal(desc, rec, level_matrices):
    """The two-level cascade of one record with primal_active_set, including qpOASES' regularisation semantics
    (solve with H + eps I, then numRegularisationSteps proximal re-solves g <- g - eps x_prev).
    Returns (x0, x1, y1)."""
    xs = []
    x0 = None
    for level in (0, 1):
        A, b, C, lA, uA, eps = level_matrices(desc, rec, level, x0)
        x, y, _ = primal_active_set(A, b, C, lA, uA, eps)
        for _ in range(desc.n_reg_steps if eps > 0.0 else 0):
            x, y, _ = primal_active_set(A, b, C, lA, uA, eps, xp=x)
        xs.append(x)
        x0 = x
    return xs[0], xs[1], y
