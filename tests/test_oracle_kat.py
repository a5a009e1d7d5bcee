"""T0/T1: analytic known-answer tests and brute-force enumeration for the oracle's dense QP core
(the Goldfarb-Idnani restatement that stands in for qpOASES)."""
import numpy as np
import pytest

from tests.qp_ref import INF, brute_force_qp, kkt_numpy

MODES = (0, 1)   # Cholesky of formed H (reference numerics) / QR of the stacked matrix


@pytest.mark.parametrize("mode", MODES)
def test_unconstrained_is_normal_equations(oracle_mod, mode):
    rng = np.random.default_rng(1)
    A, b = rng.normal(size=(12, 7)), rng.normal(size=12)
    st, x, y, it, kkt = oracle_mod.dense_qp(A, b, np.zeros((0, 7)), np.zeros(0), np.zeros(0), 1e-9, mode=mode)
    assert st == 0 and it == 0
    np.testing.assert_allclose(x, np.linalg.solve(A.T @ A + 1e-9 * np.eye(7), A.T @ b), rtol=1e-9)
    assert kkt < 1e-12


@pytest.mark.parametrize("mode", MODES)
def test_box_clamped_separable(oracle_mod, mode):
    # min 1/2 ||x - t||^2, l <= x <= u  ->  x = clip(t), y = x - t on the active side
    t = np.array([3.0, -2.0, 0.25, 10.0, -0.5])
    lo, hi = -np.ones(5), np.ones(5)
    st, x, y, it, kkt = oracle_mod.dense_qp(np.eye(5), t, np.eye(5), lo, hi, 0.0, mode=mode)
    assert st == 0
    np.testing.assert_allclose(x, np.clip(t, lo, hi), atol=1e-14)
    np.testing.assert_allclose(y, np.clip(t, lo, hi) - t, atol=1e-14)   # >0 at lower, <0 at upper
    assert it == 3 and kkt < 1e-14


@pytest.mark.parametrize("mode", MODES)
def test_single_active_halfspace(oracle_mod, mode):
    # projection of t onto {a'x <= 1}:  x = t - a (a't - 1)/||a||^2
    t, a = np.array([2.0, 1.0, -1.0]), np.array([1.0, 2.0, 2.0])
    st, x, y, it, kkt = oracle_mod.dense_qp(np.eye(3), t, a[None], np.array([-INF]), np.array([0.5]), 0.0, mode=mode)
    lam = (a @ t - 0.5) / (a @ a)
    np.testing.assert_allclose(x, t - lam * a, atol=1e-14)
    np.testing.assert_allclose(y, [-lam], atol=1e-14)
    assert st == 0 and it == 1


@pytest.mark.parametrize("mode", MODES)
def test_equality_constrained_least_norm(oracle_mod, mode):
    # H = eps I only (no task): min eps/2 ||x||^2 s.t. Cx = d -> least-norm solution (the regularisation
    # defines the answer exactly as for the cost-free contact forces, ref:src/ForceAcc.cpp:131)
    rng = np.random.default_rng(3)
    C, d = rng.normal(size=(3, 6)), rng.normal(size=3)
    st, x, y, it, kkt = oracle_mod.dense_qp(np.zeros((1, 6)), np.zeros(1), C, d, d, 2.221e-9, mode=mode)
    assert st == 0
    np.testing.assert_allclose(x, np.linalg.pinv(C) @ d, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("mode", MODES)
def test_degenerate_tie_and_redundant_equality(oracle_mod, mode):
    # two identical inequality rows active at the same point + a duplicated equality
    t = np.array([2.0, 2.0])
    C = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, 1.0]])
    lA = np.array([-INF, -INF, 0.5, 0.5])
    uA = np.array([1.0, 1.0, 0.5, 0.5])
    st, x, y, it, kkt = oracle_mod.dense_qp(np.eye(2), t, C, lA, uA, 0.0, mode=mode)
    assert st == 0
    np.testing.assert_allclose(x, [1.0, 0.5], atol=1e-14)
    assert abs(y[0] + y[1] + 1.0) < 1e-13 and kkt < 1e-13


@pytest.mark.parametrize("mode", MODES)
def test_infeasible_detected(oracle_mod, mode):
    C = np.array([[1.0, 0.0], [1.0, 0.0]])
    st, *_ = oracle_mod.dense_qp(np.eye(2), np.zeros(2), C, np.array([2.0, -INF]), np.array([INF, 1.0]), 0.0, mode=mode)
    assert st == 2
    Ceq = np.array([[1.0, 1.0], [2.0, 2.0]])
    st, *_ = oracle_mod.dense_qp(np.eye(2), np.zeros(2), Ceq, np.array([1.0, 3.0]), np.array([1.0, 3.0]), 0.0, mode=mode)
    assert st == 2


def test_max_iter_reported(oracle_mod):
    t = np.full(6, 5.0)
    st, *_ = oracle_mod.dense_qp(np.eye(6), t, np.eye(6), -np.ones(6), np.ones(6), 0.0, max_iter=3)
    assert st == 1


def test_proximal_step_semantics(oracle_mod):
    # qpOASES solveRegularisedQP: x1 = argmin 1/2 x'(H+eps I)x + (g - eps x0)'x ; closed form for diagonal H
    hdiag, b, eps = np.array([1.0, 0.0, 4.0]), np.array([1.0, 0.0, 2.0]), 1e-3
    A = np.diag(np.sqrt(hdiag))
    bb = np.where(hdiag > 0, b * np.sqrt(hdiag), 0.0)          # g = -A'bb = -hdiag*b
    g = -hdiag * b
    x0 = -g / (hdiag + eps)
    x1 = -(g - eps * x0) / (hdiag + eps)
    st, xs0, *_ = oracle_mod.dense_qp(A, bb, np.zeros((0, 3)), np.zeros(0), np.zeros(0), eps, n_reg_steps=0)
    st, xs1, *_ = oracle_mod.dense_qp(A, bb, np.zeros((0, 3)), np.zeros(0), np.zeros(0), eps, n_reg_steps=1)
    np.testing.assert_allclose(xs0, x0, rtol=1e-12)
    np.testing.assert_allclose(xs1, x1, rtol=1e-12)


@pytest.mark.parametrize("seed", range(40))
def test_brute_force_enumeration(oracle_mod, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(2, 6)); m = int(rng.integers(n, 2 * n + 2)); nc = int(rng.integers(1, 8))
    A, b = rng.normal(size=(m, n)), rng.normal(size=m) * 3
    C = rng.normal(size=(nc, n))
    mid = C @ rng.normal(size=n)                                   # guarantees feasibility
    wl, wu = rng.uniform(0.0, 1.0, nc), rng.uniform(0.0, 1.0, nc)
    lA, uA = mid - wl, mid + wu
    kind = rng.integers(0, 4, nc)
    lA[kind == 1] = -INF; uA[kind == 2] = INF
    n_eq = min(int((kind == 3).sum()), n - 1)
    eq_idx = np.nonzero(kind == 3)[0][:n_eq]
    lA[eq_idx] = uA[eq_idx] = mid[eq_idx]
    rest = np.setdiff1d(np.nonzero(kind == 3)[0], eq_idx)
    uA[rest] = mid[rest] + 0.3
    eps = 1e-6
    H, g = A.T @ A + eps * np.eye(n), -A.T @ b
    ref = brute_force_qp(H, g, C, lA, uA)
    assert ref is not None
    for mode in MODES:
        st, x, y, it, kkt = oracle_mod.dense_qp(A, b, C, lA, uA, eps, mode=mode)
        assert st == 0
        np.testing.assert_allclose(x, ref[0], rtol=1e-7, atol=1e-8)
        assert max(kkt_numpy(H, g, C, lA, uA, x, y)) < 1e-9 and kkt < 1e-9
