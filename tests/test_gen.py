"""Synthetic-state generator: determinism / shard regeneration, and physical consistency of the
rigid-body quantities it feeds the QP (M SPD, Jdot*qdot by finite differences, power balance)."""
import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, layout


@pytest.mark.parametrize("ci", (1, 2))
def test_slices_regenerate_identically(ci):
    d = CONFIGS[ci]["desc"]
    full = gen.generate(d, 96, gen.config_seed(ci))
    assert np.array_equal(full[37:71], gen.generate(d, 34, gen.config_seed(ci), start=37))
    assert np.array_equal(full, gen.generate(d, 96, gen.config_seed(ci), chunk=17))
    assert not np.array_equal(full[0], full[1])
    assert np.isfinite(full).all()


@pytest.mark.parametrize("n_a", (29, 33))
def test_mass_matrix_spd_and_jacobian_structure(n_a):
    rob = gen.robot_for(n_a)
    rng = np.random.default_rng(0)
    B = 8
    q = rob.q_home + rng.uniform(-0.3, 0.3, (B, n_a)); qd = rng.normal(0, 0.5, (B, n_a))
    R0 = gen._rpy(*rng.uniform(-0.2, 0.2, (3, B)))
    dyn = rob.dynamics(q, qd, R0, rng.normal(size=(B, 3)), rng.normal(size=(B, 3)), rng.normal(size=(B, 3)), [0] + rob.foot)
    assert (np.linalg.eigvalsh(dyn["M"]) > 1e-6).all()
    assert abs(dyn["M"][0, 0, 0] - rob.mass.sum()) < 1e-9          # total mass on the base translation block
    Jw = dyn["links"][0]["J"]
    np.testing.assert_allclose(Jw[:, :, :6], np.tile(np.eye(6), (B, 1, 1)), atol=1e-15)  # pelvis = floating base
    assert np.abs(Jw[:, :, 6:]).max() == 0.0


def _advance(R0, p0, q, v0, w0, qd, dt):
    th = np.linalg.norm(w0, axis=1)
    ax = w0 / th[:, None]
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 2], ax[:, 1], ax[:, 2], -ax[:, 0], -ax[:, 1], ax[:, 0]
    a = (th * dt)[:, None, None]
    E = np.eye(3)[None] + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)
    return E @ R0, p0 + dt * v0, q + dt * qd


def test_jdot_qdot_and_power_balance():
    rob = gen.robot_for(29)
    rng = np.random.default_rng(5)
    B = 6
    q = rob.q_home + rng.uniform(-0.3, 0.3, (B, 29)); qd = rng.normal(0, 0.5, (B, 29))
    R0 = gen._rpy(*rng.uniform(-0.3, 0.3, (3, B)))
    p0, v0, w0 = rng.normal(size=(B, 3)), rng.normal(0, 0.3, (B, 3)), rng.normal(0, 0.3, (B, 3))
    links = [0] + rob.foot + rob.hand
    v = np.concatenate([v0, w0, qd], axis=1)
    d0 = rob.dynamics(q, qd, R0, p0, v0, w0, links)
    dt = 1e-6
    dp = rob.dynamics(*(lambda r: (r[2], qd, r[0], r[1], v0, w0, links))(_advance(R0, p0, q, v0, w0, qd, dt)))
    dm = rob.dynamics(*(lambda r: (r[2], qd, r[0], r[1], v0, w0, links))(_advance(R0, p0, q, v0, w0, qd, -dt)))
    for b in links:                       # Jdot*qdot = d/dt (J) v  at constant generalised velocity
        fd = np.einsum("bij,bj->bi", (dp["links"][b]["J"] - dm["links"][b]["J"]) / (2 * dt), v)
        np.testing.assert_allclose(d0["links"][b]["Jdqd"], fd, atol=2e-6)
    # power balance: v'(h - G) = 1/2 v' Mdot v   (passivity of the Coriolis terms)
    zero = np.zeros_like
    G = rob.dynamics(q, zero(qd), R0, p0, zero(v0), zero(w0), [])["h"]
    Mdot = (dp["M"] - dm["M"]) / (2 * dt)
    lhs = np.einsum("bi,bi->b", v, d0["h"] - G)
    rhs = 0.5 * np.einsum("bi,bij,bj->b", v, Mdot, v)
    np.testing.assert_allclose(lhs, rhs, atol=5e-6 * max(1.0, np.abs(rhs).max()))
    # gravity term = gradient of the potential energy along any generalised velocity
    def potential(Rb, pb, qq):
        dd = rob.dynamics(qq, zero(qd), Rb, pb, zero(v0), zero(w0), list(range(rob.n_b)))
        return sum(rob.mass[i] * gen.GRAVITY * (dd["links"][i]["p"][:, 2] + np.einsum("bij,j->bi", dd["links"][i]["R"], rob.com[i])[:, 2])
                   for i in range(rob.n_b))
    Vp = potential(*_advance(R0, p0, q, v0, w0, qd, dt)); Vm = potential(*_advance(R0, p0, q, v0, w0, qd, -dt))
    np.testing.assert_allclose(np.einsum("bi,bi->b", v, G), (Vp - Vm) / (2 * dt), rtol=1e-6, atol=1e-5)
