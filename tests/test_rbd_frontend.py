"""SURVEY.md 8(f) row 1 — the on-device rigid-body front end (compact states -> records) against its CPU statement
(qppvm_b200/gen.py: Robot.dynamics + records_from_states, itself validated by finite differences in test_gen.py),
and the states -> torques path against the oracle."""
import ctypes as C

import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, layout
from tests.helpers import PRIMAL_TOL, rel_inf


def test_state_layout_agrees_with_c_abi():
    from qppvm_b200 import api, build
    build.build()
    lib = api.load_library()
    for ci in (0, 1, 2):
        d = CONFIGS[ci]["desc"]
        assert lib.qppvm_state_doubles(C.byref(api.cdesc(d))) == gen.state_doubles(d)
        assert gen.generate_states(d, 3, 1).shape == (3, gen.state_doubles(d))
        np.testing.assert_array_equal(gen.records_from_states(d, gen.generate_states(d, 3, 1)), gen.generate(d, 3, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("ci", (1, 0, 2))
def test_device_records_match_cpu_dynamics(ci):
    import torch
    from qppvm_b200 import api
    desc = CONFIGS[ci]["desc"]
    L = layout(desc)
    rob = gen.robot_for(desc.n_a)
    s = api.Solver(desc)
    s.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
    states = gen.generate_states(desc, 700, gen.config_seed(ci))
    ref = gen.records_from_states(desc, states)
    got = s.records_from_states(torch.from_numpy(states).cuda()).cpu().numpy()
    for name, a, b in (("J_waist", L.off_jwaist, L.off_jc), ("J_c", L.off_jc, L.off_M), ("M", L.off_M, L.off_h),
                       ("h", L.off_h, L.off_jdqd), ("Jdqd", L.off_jdqd, L.off_rhs), ("rest", L.off_rhs, L.rec_doubles)):
        scale = max(1.0, np.abs(ref[:, a:b]).max())
        assert np.abs(got[:, a:b] - ref[:, a:b]).max() <= 1e-11 * scale, name


@pytest.mark.gpu
def test_states_to_torques_matches_oracle(oracle_mod):
    from qppvm_b200 import api
    desc = CONFIGS[1]["desc"]
    L = layout(desc)
    rob = gen.robot_for(desc.n_a)
    s = api.Solver(desc)
    s.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
    states = gen.generate_states(desc, 3000, 99)                       # > 2 host chunks, ragged tail
    out = api.split_out(L, s.solve_states_host(states))
    o = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, gen.records_from_states(desc, states))[0])
    assert (out["status"] == 0).all() and (o["status"] == 0).all()
    assert rel_inf(out["x"], o["x"]).max() <= PRIMAL_TOL and rel_inf(out["tau"], o["tau"]).max() <= PRIMAL_TOL
    assert np.array_equal(out["active"], o["active"]) and out["kkt"].max() <= 1e-6
    with pytest.raises(api.QPError):                                   # front end needs the robot tables first
        api.Solver(desc).solve_states_host(states[:4])


@pytest.mark.gpu
def test_config3_one_million_states(oracle_mod):
    """BASELINE configs[3] at full size on one GPU: 2^20 WALK-MAN-like states (4 contacts, cones + tau-limits).
    States come from the seeded generator, records from the device front end (18.6 GB, never leave the GPU); every
    solve must converge with the in-kernel KKT certificate <= 1e-6, and a random sample is checked against the oracle."""
    import torch
    from qppvm_b200 import api
    desc, B = CONFIGS[3]["desc"], 1 << 20
    L = layout(desc)
    rob = gen.robot_for(desc.n_a)
    s = api.Solver(desc)
    s.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
    states = gen.generate_states(desc, B, gen.config_seed(3))
    recs = s.records_from_states(torch.from_numpy(states).cuda())
    out, _ = s.solve_batch(recs)
    torch.cuda.synchronize()
    tr = out[:, L.n_x + L.n_a:].contiguous().view(torch.int32)          # trailer words: status, iters, mask x4, kkt x2
    bad = torch.nonzero(tr[:, 0] != 0)[:, 0]
    # random states with scaled-down torque limits are, very rarely, genuinely infeasible (seed 20263118: one in
    # 2^20): every non-OK status must be confirmed by the oracle on the same record
    assert bad.numel() <= B // 100000
    if bad.numel():
        ob = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, recs[bad].cpu().numpy())[0])
        assert np.array_equal(ob["status"], tr[bad, 0].cpu().numpy())
    kkt = tr[:, 6:8].contiguous().view(torch.float32)
    assert float(kkt[tr[:, 0] == 0].max()) <= 1e-6
    idx = torch.from_numpy(np.random.default_rng(3).choice(B, 384, replace=False)).cuda()
    idx = idx[tr[idx, 0] == 0]
    g = api.split_out(L, out[idx].cpu().numpy())
    o = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, recs[idx].cpu().numpy())[0])
    assert rel_inf(g["x"], o["x"]).max() <= PRIMAL_TOL and np.array_equal(g["active"], o["active"])
