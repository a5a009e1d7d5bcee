"""SURVEY.md 8(f) row 3 — the command side: integration of the solved accelerations
(ref:src/ForceAcc.cpp:225-226) and closed-loop rollouts (front end -> 2-level solve -> integrate).
CPU tests pin the numpy statement (gen.integrate_states) to closed forms and run it in closed loop with the
oracle; GPU tests compare integrate_states_kernel / qppvm_rollout_states with that statement."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, layout

DT = 1e-3


def _fake_out(desc, B, rng, status=None):
    L = layout(desc)
    out = np.zeros((B, L.out_doubles))
    out[:, :L.n_x] = rng.normal(0, 3.0, (B, L.n_x))
    tr = np.zeros((B, 8), dtype=np.int32)
    if status is not None:
        tr[:, 0] = status
    out[:, L.n_x + desc.n_a:L.n_x + desc.n_a + 4] = tr.view(np.float64)
    return out


@pytest.mark.parametrize("ci", (1, 2))
def test_integration_closed_forms(ci):
    desc = CONFIGS[ci]["desc"]
    L, o = layout(desc), gen.state_offsets(desc)
    rng = np.random.default_rng(3)
    B = 64
    st = gen.generate_states(desc, B, 17)
    out = _fake_out(desc, B, rng)
    new = gen.integrate_states(desc, st, out, DT)
    sl = lambda a, k: a[:, o[k][0]:o[k][1]]
    qdd, a0, al0 = out[:, 6:L.n_v], out[:, :3], out[:, 3:6]
    np.testing.assert_allclose(sl(new, "q"), sl(st, "q") + DT * sl(st, "qd") + 0.5 * DT * DT * qdd, rtol=0, atol=1e-15)
    np.testing.assert_allclose(sl(new, "qd"), sl(st, "qd") + DT * qdd, rtol=0, atol=1e-15)
    np.testing.assert_allclose(sl(new, "p0"), sl(st, "p0") + DT * sl(st, "tw")[:, :3] + 0.5 * DT * DT * a0, rtol=0, atol=1e-15)
    np.testing.assert_allclose(sl(new, "tw"), sl(st, "tw") + DT * out[:, :6], rtol=0, atol=1e-15)
    th = DT * sl(st, "tw")[:, 3:] + 0.5 * DT * DT * al0
    Rref = Rotation.from_rotvec(th).as_matrix() @ sl(st, "R0").reshape(B, 3, 3)
    Rn = sl(new, "R0").reshape(B, 3, 3)
    np.testing.assert_allclose(Rn, Rref, rtol=0, atol=1e-14)
    np.testing.assert_allclose(Rn @ Rn.transpose(0, 2, 1), np.tile(np.eye(3), (B, 1, 1)), rtol=0, atol=1e-14)
    for k in ("gains", "ori_err", "foot_err", "mu", "tau_scale"):       # references and parameters are not integrated
        np.testing.assert_array_equal(sl(new, k), sl(st, k))
    # large rotation step and the zero step go through the two branches of the exponential
    big = st.copy(); big[:, o["tw"][0] + 3:o["tw"][1]] = rng.normal(0, 300.0, (B, 3))
    Rb = gen.integrate_states(desc, big, out, DT)[:, o["R0"][0]:o["R0"][1]].reshape(B, 3, 3)
    thb = DT * big[:, o["tw"][0] + 3:o["tw"][1]] + 0.5 * DT * DT * al0
    np.testing.assert_allclose(Rb, Rotation.from_rotvec(thb).as_matrix() @ sl(st, "R0").reshape(B, 3, 3), rtol=0, atol=1e-13)
    still = st.copy(); still[:, o["tw"][0]:o["tw"][1]] = 0.0; still[:, o["qd"][0]:o["qd"][1]] = 0.0
    zero = _fake_out(desc, B, rng); zero[:, :L.n_x] = 0.0
    np.testing.assert_array_equal(gen.integrate_states(desc, still, zero, DT), still)


def test_failed_solves_leave_their_state_untouched():
    desc = CONFIGS[1]["desc"]
    rng = np.random.default_rng(5)
    st = gen.generate_states(desc, 16, 2)
    status = np.zeros(16, dtype=np.int32); status[::3] = 2; status[1::5] = 1
    new = gen.integrate_states(desc, st, _fake_out(desc, 16, rng, status), DT)
    bad = status != 0
    np.testing.assert_array_equal(new[bad], st[bad])
    assert (np.abs(new[~bad] - st[~bad]).max(axis=1) > 0).all()


def test_closed_loop_with_oracle_tracks_the_waist_reference(oracle_mod):
    """Closed loop on the CPU statement over 150 periods.  The references are captured once (ref:src/ForceAcc.cpp:158-164,
    waist position reference = initial - 0.1 z, :181), so the stored waist error must SHRINK as the waist moves down --
    with a constant error the waist would be driven down at lambda e / lambda2 = 0.5 m/s for ever."""
    desc = CONFIGS[1]["desc"]
    o = gen.state_offsets(desc)
    st = gen.generate_states(desc, 24, 11)
    st[:, o["qd"][0]:o["qd"][1]] *= 0.1                       # start near rest: the test is about the waist task
    st[:, o["tw"][0]:o["tw"][1]] *= 0.1
    p_start = st[:, o["p0"][0]:o["p0"][1]].copy()
    e_start = st[:, o["waist_pos_err"][0]:o["waist_pos_err"][1]].copy()
    assert np.array_equal(e_start, np.tile([0.0, 0.0, -0.1], (24, 1)))
    for _ in range(150):
        recs = gen.records_from_states(desc, st)
        out, _ = oracle_mod.solve_batch(desc, recs)
        g = oracle_mod.split_out(desc, out)
        assert (g["status"] == 0).mean() >= 0.9
        st = gen.integrate_states(desc, st, out.view(np.float64).reshape(len(st), -1), DT, recs=recs)
    # the stored error is reference - current: what is left of it plus the distance travelled is the initial error
    moved = st[:, o["p0"][0]:o["p0"][1]] - p_start
    e_now = st[:, o["waist_pos_err"][0]:o["waist_pos_err"][1]]
    np.testing.assert_allclose(e_now + moved, e_start, rtol=0, atol=2e-3)
    ok = oracle_mod.split_out(desc, out)["status"] == 0
    # critically damped second-order response, omega = 10 sqrt(gain): e(0.15 s) = 0.1 (1 + 1.5) exp(-1.5) = 0.056
    assert (np.abs(e_now[ok, 2]) < 0.07).all() and (moved[ok, 2] < -0.03).all()
    R = st[:, o["R0"][0]:o["R0"][1]].reshape(-1, 3, 3)
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.tile(np.eye(3), (len(st), 1, 1)), rtol=0, atol=1e-12)
    assert np.isfinite(st).all()
    assert np.abs(st[:, o["qd"][0]:o["qd"][1]]).max() < 20.0


# ---------------------------------------------------------------------------------------------- GPU

def _solver(desc):
    from qppvm_b200 import api
    rob = gen.robot_for(desc.n_a)
    s = api.Solver(desc)
    s.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("ci", (1, 2))
def test_integrate_kernel_matches_numpy(ci):
    import torch
    desc = CONFIGS[ci]["desc"]
    rng = np.random.default_rng(8)
    B = 777
    st = gen.generate_states(desc, B, 23)
    status = np.zeros(B, dtype=np.int32); status[::7] = 2
    out = _fake_out(desc, B, rng, status)
    s = _solver(desc)
    d = torch.from_numpy(st).cuda()
    s.integrate_states(d, torch.from_numpy(out).cuda(), DT)
    torch.cuda.synchronize()
    ref = gen.integrate_states(desc, st, out, DT)
    np.testing.assert_allclose(d.cpu().numpy(), ref, rtol=0, atol=1e-14)
    np.testing.assert_array_equal(d.cpu().numpy()[::7], st[::7])
    # with the tick's records: the stored task errors follow the motion
    recs = gen.records_from_states(desc, st)
    d = torch.from_numpy(st).cuda()
    s.integrate_states(d, torch.from_numpy(out).cuda(), DT, records=torch.from_numpy(recs).cuda())
    torch.cuda.synchronize()
    ref = gen.integrate_states(desc, st, out, DT, recs=recs)
    o = gen.state_offsets(desc)
    assert np.abs(ref[:, o["waist_pos_err"][0]:o["waist_pos_err"][1]] - st[:, o["waist_pos_err"][0]:o["waist_pos_err"][1]]).max() > 0
    np.testing.assert_allclose(d.cpu().numpy(), ref, rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_rollout_matches_cpu_closed_loop(oracle_mod):
    import torch
    from qppvm_b200 import api
    from tests.helpers import rel_inf
    desc = CONFIGS[1]["desc"]
    L = layout(desc)
    B, T = 256, 6
    st0 = gen.generate_states(desc, B, 31)
    s = _solver(desc)
    d = torch.from_numpy(st0).cuda()
    out = s.rollout_states(d, T, DT)
    torch.cuda.synchronize()
    st = st0
    for _ in range(T):
        o_out, _ = oracle_mod.solve_batch(desc, gen.records_from_states(desc, st))
        o_out = o_out.view(np.float64).reshape(B, -1)
        st_prev, st = st, gen.integrate_states(desc, st, o_out, DT, recs=gen.records_from_states(desc, st))
    g, o = api.split_out(L, out.cpu().numpy()), oracle_mod.split_out(desc, o_out)
    assert (g["status"] == 0).all() and np.array_equal(g["status"], o["status"])
    assert rel_inf(g["x"], o["x"]).max() <= 1e-6 and np.array_equal(g["active"], o["active"])
    np.testing.assert_allclose(d.cpu().numpy(), st, rtol=0, atol=1e-9)
    # one rollout of T ticks == T single-tick rollouts == front end + solve + integrate called by hand
    d2 = torch.from_numpy(st0).cuda()
    for _ in range(T):
        out2 = s.rollout_states(d2, 1, DT)
    d3 = torch.from_numpy(st0).cuda()
    for _ in range(T):
        r3 = s.records_from_states(d3)
        out3, _ = s.solve_batch(r3)
        s.integrate_states(d3, out3, DT, records=r3)
    torch.cuda.synchronize()
    # the two cold-started forms agree to the bit; the T-tick rollout starts every tick but the first from the working
    # sets of the previous one (the hot start of the reference's persistent solver): same minimiser, another pivot
    # order -- equal within the solver's accuracy, same active set, fewer working-set changes
    assert torch.equal(d3, d2) and torch.equal(out3, out2)
    g2 = api.split_out(L, out2.cpu().numpy())
    assert rel_inf(g["x"], g2["x"]).max() <= 1e-8 and np.array_equal(g["active"], g2["active"])
    np.testing.assert_allclose(d.cpu().numpy(), d2.cpu().numpy(), rtol=0, atol=1e-10)
    assert (g["iters0"] + g["iters1"]).mean() < (g2["iters0"] + g2["iters1"]).mean()


@pytest.mark.gpu
def test_rollout_chunks_are_independent():
    """More states than one round of rollout lanes (4 x 16 384): same bits as the two parts rolled out separately."""
    import torch
    desc = CONFIGS[1]["desc"]
    B = 65536 + 3000
    st0 = torch.from_numpy(gen.generate_states(desc, B, 41)).cuda()
    s = _solver(desc)
    whole = st0.clone()
    out = s.rollout_states(whole, 2, DT)
    a, b = st0[:65536].clone(), st0[65536:].clone()
    oa, ob = s.rollout_states(a, 2, DT).clone(), s.rollout_states(b, 2, DT).clone()
    torch.cuda.synchronize()
    assert torch.equal(whole, torch.cat([a, b])) and torch.equal(out, torch.cat([oa, ob]))
    empty = st0[:0].clone()
    assert s.rollout_states(empty, 3, DT).shape[0] == 0
