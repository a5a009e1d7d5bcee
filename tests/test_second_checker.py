"""Parity cannot be pinned (no qpOASES / OpenSoT here, no upstream goldens), so the oracle is cross-examined by solvers
that share nothing with it:

* a PRIMAL active-set method (tests/qp_ref.py: feasible start from HiGHS' LP phase 1, blocking constraints enter, wrong-
  signed multipliers leave; the working set is re-factorised from scratch each iteration) -- the oracle and the CUDA
  kernels are DUAL active-set methods, so agreement on the strongly active set is not agreement by construction;
* HiGHS' QP solver on hundreds of records.

Every level QP is strictly convex, so the minimiser is unique: the solvers must agree on it and on the constraints with a
non-zero multiplier.  Level 1 is posed with the oracle's level-0 solution (the same QP for both solvers)."""
import dataclasses

import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, Desc, FLAG_COM_TASK, layout
from tests.assemble_np import level_matrices
from tests.helpers import strongly_active
from tests.qp_ref import highs_qp, primal_active_set


# 10: the CoM force task at level 1 (cones + torque limits).  (The Torque-kind variants are certified by the independent
# numpy KKT check of test_oracle_crosscheck.py: their level 1 sits on degenerate vertices, where multipliers are not unique.)
EXTRA = {10: Desc(n_a=29, n_contacts=2, flags=FLAG_COM_TASK | 3)}


@pytest.mark.parametrize("ci,count", [(1, 1000), (2, 1000), (0, 500), (10, 200)])
def test_primal_active_set_finds_the_same_point_and_active_set(oracle_mod, ci, count):
    desc = EXTRA[ci] if ci in EXTRA else CONFIGS[ci]["desc"]
    L = layout(desc)
    n, nr = L.n_x, L.n_rows
    recs = gen.generate(desc, count, gen.config_seed(ci) + 17)
    out, dg = oracle_mod.solve_batch(desc, recs, diag=True)
    o = oracle_mod.split_out(desc, out)
    assert (o["status"] == 0).all()
    rel1 = np.zeros(count)
    n_degenerate = 0
    for i in range(count):
        x0o, eo = dg[i, :n], dg[i, n + 2 * nr:n + 2 * nr + 6]
        for lev in (0, 1):
            A, b, C, lA, uA, eps = level_matrices(desc, recs[i], lev, x0o)
            x, y, _ = primal_active_set(A, b, C, lA, uA, eps)
            for _ in range(desc.n_reg_steps):                 # qpOASES' proximal re-solve
                x, y, _ = primal_active_set(A, b, C, lA, uA, eps, xp=x)
            yo = dg[i, n + lev * nr:n + (lev + 1) * nr][:len(y)]
            xo = x0o if lev == 0 else o["x"][i]
            assert np.array_equal(strongly_active(y[None])[0], strongly_active(yo[None])[0]), (i, lev)
            f = lambda v: 0.5 * np.sum((A @ v - b) ** 2) + 0.5 * eps * v @ v
            assert abs(f(x) - f(xo)) <= 1e-9 * max(1.0, abs(f(xo))), (i, lev)
            if lev == 0:
                assert np.abs(A @ x - eo).max() <= 1e-9 * max(1.0, np.abs(eo).max()), i     # what level 1 inherits
                n_degenerate += int((np.abs(yo[6:]) > 0).any())                           # level 0 limited by a constraint
            else:
                rel1[i] = np.abs(x - xo).max() / max(1.0, np.abs(xo).max())
                nv = L.n_v                                     # the accelerations are pinned by the postural task: always 1e-6
                assert np.abs(x[:nv] - xo[:nv]).max() <= 1e-6 * max(1.0, np.abs(xo[:nv]).max()), i
    # north_star's 1e-6 on the final point.  The few per cent of records above it differ along the INTERNAL force of the
    # contacts (e.g. squeezing along the line between two point contacts: the dynamics do not see it, no task weighs
    # it), which only the 2.2e-9 regularisation defines: a gradient rounding error of 1e-12 moves the point by
    # 1e-12 / 2.2e-9 there.  Same objective to 1e-9, same active set, both KKT-exact, accelerations equal to 1e-6 on EVERY record (above).
    assert (rel1 <= 1e-6).mean() >= 0.95 and rel1.max() <= 1e-3, (rel1.max(), (rel1 > 1e-6).sum())
    if ci != 1:
        assert n_degenerate >= 0.2 * count                     # the sample does contain bound-limited level-0 optima


@pytest.mark.parametrize("ci", (1, 0, 2))
def test_highs_on_hundreds_of_records(oracle_mod, ci):
    """HiGHS stops at a KKT residual of ~1e-6, i.e. an x error of ~1e-6 / lambda_min(H): on the well-conditioned variant
    (eps = 1) the two solvers must agree to 1e-6 relative on every record; at the reference's eps HiGHS bounds the oracle's
    objective from above."""
    desc = CONFIGS[ci]["desc"]
    d2 = dataclasses.replace(desc, eps_regularisation=1.0 / 2.221e-13, n_reg_steps=0)   # eps = 1
    L = layout(d2)
    n = L.n_x
    count = 120
    recs = gen.generate(d2, count, gen.config_seed(ci) + 29)
    for dd, exact in ((d2, True), (desc, False)):
        out, dg = oracle_mod.solve_batch(dd, recs, diag=True)
        o = oracle_mod.split_out(dd, out)
        assert (o["status"] == 0).all()
        solved, errs = 0, []
        for i in range(count if exact else 40):
            x0 = dg[i, :n]
            for lev, x in ((0, x0), (1, o["x"][i])):
                A, b, C, lA, uA, eps = level_matrices(dd, recs[i], lev, x0)
                H, g = A.T @ A + eps * np.eye(n), -A.T @ b
                ok, xh, _ = highs_qp(H, g, C, lA, uA)
                if exact and ok:
                    errs.append(np.abs(xh - x).max() / max(1.0, np.abs(x).max()))
                elif ok:
                    f = lambda v: 0.5 * v @ H @ v + g @ v
                    assert f(x) <= f(xh) + 1e-7 * max(1.0, abs(f(xh))), (i, lev)
                solved += int(ok)
        assert solved >= (0.97 * 2 * count if exact else 40)     # (HiGHS gives up on a few level-1 problems)
        if exact:                                              # HiGHS' own stopping tolerance shows on a few per cent of the solves
            errs = np.array(errs)
            assert (errs < 1e-6).mean() >= 0.97 and errs.max() < 1e-4, (errs.max(), (errs >= 1e-6).sum())
