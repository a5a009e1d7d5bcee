"""T4: the plugin drop-ins behind a fake XBot::Handle / RobotInterface / ModelInterface.
CPU part: the libraries build and export the reference's registration symbols.  GPU part: the harness
(qppvm_b200/plugin/plugin_test.cpp) loads each plugin like XBotCore would and plays recorded synthetic states
through init_control_plugin / on_start / control_loop / close; the record the plugin packed is compared with an
independent numpy packing of the same state, and the commanded torques with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import Desc, KIND_TORQUE, FLAG_COM_TASK, FLAG_ELBOW_TASKS, FLAG_JOINT_LIMITS, layout
from tests.helpers import PRIMAL_TOL, rel_inf

PLUG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qppvm_b200", "plugin")


@pytest.fixture(scope="module")
def built():
    from qppvm_b200 import build as native
    native.build()
    from qppvm_b200.plugin import build as pb
    pb.build()
    return PLUG


def test_plugin_libraries_export_registration_symbols(built):
    # The plugin libraries resolve XBot symbols against the hosting process (libXBotInterface in production, the
    # harness here), so the exported registration symbols are read from the dynamic symbol table.
    def exported(lib):
        out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(built, lib)], capture_output=True, text=True, check=True).stdout
        return {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    # REGISTER_XBOT_PLUGIN_(XBotPlugin::ForceAccExample)  ref:src/ForceAcc.cpp:26 ; library names ref:CMakeLists.txt:48-49
    assert {"create_instance", "destroy_instance"} <= exported("libForceAccPlugin.so")
    # REGISTER_XBOT_PLUGIN(QPPVMPlugin, demo::QPPVMPlugin)  ref:src/QPPVMPlugin.cpp:29
    assert "QPPVMPlugin_factory" in exported("libQPPVMPlugin.so")
    assert os.access(os.path.join(built, "plugin_test"), os.X_OK)


def _quat(R):
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s; q[j] = (R[j, i] + R[i, j]) / s; q[k] = (R[k, i] + R[i, k]) / s; q[3] = (R[k, j] - R[j, k]) / s
    return q


def _ori_err(Rd, R):
    qd, q = _quat(Rd), _quat(R)
    if qd @ q < 0:
        qd = -qd
    return q[3] * qd[:3] - qd[3] * q[:3] - np.cross(qd[:3], q[:3])


def _states(rob, T, seed, floating):
    rng = np.random.default_rng(seed)
    na = rob.n_a
    q = rob.q_home[None] + rng.uniform(-0.15, 0.15, (T, na)); qd = rng.normal(0, 0.3, (T, na))
    if floating:
        rpy = rng.uniform(-0.1, 0.1, (T, 3)); R0 = gen._rpy(rpy[:, 0], rpy[:, 1], rpy[:, 2])
        p0 = np.array([0, 0, 0.6]) + rng.uniform(-0.02, 0.02, (T, 3)); v0 = rng.normal(0, 0.1, (T, 3)); w0 = rng.normal(0, 0.1, (T, 3))
    else:
        rpy = np.zeros((T, 3)); R0 = np.tile(np.eye(3), (T, 1, 1)); p0 = v0 = w0 = np.zeros((T, 3))
    return q, qd, rpy, R0, p0, v0, w0


def _write_states(path, T, nv, per_tick):
    with open(path, "wb") as f:
        for t in range(T):
            for a in per_tick(t):
                f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())


def _run(built, lib, factory, states, out, T, nv, floating, links, env=None):
    cmd = [os.path.join(built, "plugin_test"), os.path.join(built, lib), factory, states, out, str(T), str(nv), str(int(floating))] + links
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr + r.stdout
    return r.stdout, r.stderr


@pytest.mark.gpu
# 39: the reference's literal shape, n_v = 45 (ref:src/ForceAcc.cpp:58-70); "com": FORCEACC_PLUGIN_STACK=com stacks the
# _com_task the reference only constructs (ref:src/ForceAcc.cpp:103)
@pytest.mark.parametrize("n_a,stack", ((33, ""), (39, ""), (39, "com")))
def test_forceacc_plugin_boundary(built, oracle_mod, tmp_path, n_a, stack):
    rob = gen.robot_for(n_a)
    nv, T = n_a + 6, 10
    desc = Desc(n_a=n_a, n_contacts=4, flags=FLAG_COM_TASK if stack == "com" else 0)
    L = layout(desc)
    links = ["pelvis", "foot_fl", "foot_fr", "foot_hr", "foot_hl"]
    bodies = [0] + rob.foot + rob.hand
    q, qd, rpy, R0, p0, v0, w0 = _states(rob, T, 5, True)
    dyn = rob.dynamics(q, qd, R0, p0, v0, w0, bodies)
    home = np.concatenate([np.zeros(6), rob.q_home])
    qfull = np.concatenate([p0, rpy, q], axis=1); vfull = np.concatenate([v0, w0, qd], axis=1)
    M = dyn["M"].copy(); h = dyn["h"].copy()
    M[T - 1, 3, 3] = np.nan                                    # last tick: non-finite model -> solver failure path

    def tick(t):
        yield from (qfull[t], vfull[t], home, M[t], h[t], np.ones(nv), p0[t], v0[t], R0[t], w0[t])
        for b in bodies:
            lk = dyn["links"][b]
            yield from (lk["J"][t], lk["Jdqd"][t], lk["R"][t], lk["p"][t], lk["J"][t] @ vfull[t])
    _write_states(tmp_path / "s.bin", T, nv, tick)
    os.environ["QPPVM_TRACE_DIR"] = str(tmp_path)
    stdout, stderr = _run(built, "libForceAccPlugin.so", "create_instance", str(tmp_path / "s.bin"), str(tmp_path / "o.bin"), T, nv, True, links,
                          env={"FORCEACC_PLUGIN_STACK": stack})
    per = 3 + nv + L.rec_doubles + L.out_doubles
    o = np.fromfile(tmp_path / "o.bin").reshape(T, per)
    status, moved, nerr = o[:, 0], o[:, 1], o[:, 2]
    eff, rec, out = o[:, 3:3 + nv], o[:, 3 + nv:3 + nv + L.rec_doubles], o[:, 3 + nv + L.rec_doubles:]
    # ---- failure convention (ref:src/ForceAcc.cpp:189-193): error logged, early return, nothing commanded
    assert (status[:-1] == 0).all() and status[-1] != 0
    assert (moved[:-1] == 1).all() and moved[-1] == 0 and nerr[-1] == 1 and nerr[-2] == 0
    assert "Unable to solve!!!" in stderr and "close=1" in stdout
    # ---- the record the plugin packed == independent packing of the same state (lambda = 100, lambda2 = 20)
    exp = np.zeros((T, L.rec_doubles))
    ref_pose = {b: (dyn["links"][b]["R"][0], dyn["links"][b]["p"][0]) for b in bodies}
    for t in range(T):
        for i, b in enumerate(bodies):
            lk = dyn["links"][b]
            Rr, pr = ref_pose[b]
            if b == 0:
                pr = dyn["links"][0]["p"][0] - np.array([0, 0, 0.1])              # ref:src/ForceAcc.cpp:181
            e = np.concatenate([pr - lk["p"][t], _ori_err(Rr, lk["R"][t])])
            rhs = 100.0 * e - 20.0 * (lk["J"][t] @ vfull[t])
            if b == 0:
                exp[t, L.off_jwaist:L.off_jwaist + 6 * nv] = lk["J"][t].ravel()
                exp[t, L.off_rhs:L.off_rhs + 6] = rhs; exp[t, L.off_jdqd:L.off_jdqd + 6] = lk["Jdqd"][t]
            else:
                c = i - 1
                exp[t, L.off_jc + c * 6 * nv:L.off_jc + (c + 1) * 6 * nv] = lk["J"][t].ravel()
                exp[t, L.off_rhs + 6 * (1 + c):L.off_rhs + 6 * (2 + c)] = rhs
                exp[t, L.off_jdqd + 6 * (1 + c):L.off_jdqd + 6 * (2 + c)] = lk["Jdqd"][t]
                exp[t, L.off_fbox + 6 * c:L.off_fbox + 6 * c + 6] = [-1000, -1000, 10, 1000, 1000, 1000]
        exp[t, L.off_rhs + 30:L.off_rhs + 30 + nv] = 100.0 * (home - qfull[t]) - 20.0 * vfull[t]
        exp[t, L.off_M:L.off_M + nv * (nv + 1) // 2] = gen.pack_lower(M[t]); exp[t, L.off_h:L.off_h + nv] = h[t]
        if stack == "com":
            # sum f_i = m (100 (c_ref - c) - 20 cdot) + m g z, sum (p_i - c) x f_i = -20 L_c, with m, c, momentum from M
            def com_of(tt):
                m = M[tt, 0, 0]
                return m, np.array([M[tt, 1, 5], M[tt, 2, 3], M[tt, 0, 4]]) / m
            (m, dcm), (_, dcm0) = com_of(t), com_of(0)
            cpos, cref = dyn["links"][0]["p"][t] + dcm, dyn["links"][0]["p"][0] + dcm0
            mom = M[t, :6] @ vfull[t]
            A = np.zeros((6, 12))
            for ci, b in enumerate(bodies[1:]):
                r = dyn["links"][b]["p"][t] - cpos
                A[:3, 3 * ci:3 * ci + 3] = np.eye(3)
                A[3:, 3 * ci:3 * ci + 3] = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
            bl = m * (100.0 * (cref - cpos) - 20.0 * mom[:3] / m) + np.array([0, 0, m * 9.81])
            exp[t, L.off_com:L.off_com + 72] = A.ravel()
            exp[t, L.off_com + 72:L.off_com + 78] = np.concatenate([bl, -20.0 * (mom[3:] - np.cross(dcm, mom[:3]))])
    np.testing.assert_allclose(rec[:-1], exp[:-1], rtol=1e-12, atol=1e-9)
    # ---- what was commanded == the oracle's answer for that record
    oo = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, rec[:-1])[0])
    assert (oo["status"] == 0).all()
    assert rel_inf(out[:-1, :L.n_x], oo["x"]).max() <= PRIMAL_TOL
    assert rel_inf(eff[:-1, 6:], oo["tau"]).max() <= PRIMAL_TOL and np.abs(eff[:-1, :6]).max() == 0.0
    # ---- MatLogger-compatible trace (ref:src/ForceAcc.cpp:200,233-236, flushed in close(): ForceAcc.h:43): one column per
    # commanded tick under the reference's variable names, readable as a MAT-file
    import scipy.io
    tr = scipy.io.loadmat(str(tmp_path / "opensot_force_acc_example.mat"))
    assert tr["x"].shape == (L.n_x, T - 1) and tr["tau"].shape == (nv, T - 1) and tr["foot_fl_wrench"].shape == (6, T - 1)
    np.testing.assert_array_equal(tr["x"].T, out[:-1, :L.n_x])
    np.testing.assert_array_equal(tr["tau"].T[:, 6:], out[:-1, L.n_x:L.n_x + L.n_a])
    np.testing.assert_array_equal(tr["qddot_value"].T, out[:-1, :nv])
    assert np.abs(tr["dyn_feas_residual"]).max() <= 1e-8          # what checkConstraint(_x) reports at ForceAcc.cpp:203
    assert (tr["qp_status_iters_kkt"][0] == 0).all() and tr["qp_status_iters_kkt"][3].max() <= 1e-6


@pytest.mark.gpu
def test_qppvm_plugin_boundary(built, oracle_mod, tmp_path):
    rob = gen.robot_for(29)
    n, T = 29, 8
    desc = Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=0, eps_regularisation=1.0)
    L = layout(desc)
    links = ["arm1_7", "arm2_7"]                                # left, right (ref:src/QPPVMPlugin.cpp:132,145)
    bodies = [rob.hand[0], rob.hand[1]]
    q, qd, rpy, R0, p0, v0, w0 = _states(rob, T, 9, False)
    dyn = rob.dynamics(q, qd, R0, p0, v0, w0, bodies)
    M = dyn["M"][:, 6:, 6:].copy(); h = dyn["h"][:, 6:].copy()
    tmax = rob.tau_max * 0.6
    M[T - 1, 2, 2] = np.nan

    def tick(t):
        yield from (q[t], qd[t], rob.q_home, M[t], h[t], tmax, np.zeros(3), np.zeros(3), np.eye(3), np.zeros(3))
        for b in bodies:
            lk = dyn["links"][b]
            yield from (lk["J"][t][:, 6:], lk["Jdqd"][t], lk["R"][t], lk["p"][t], lk["J"][t][:, 6:] @ qd[t])
    _write_states(tmp_path / "s.bin", T, n, tick)
    stdout, stderr = _run(built, "libQPPVMPlugin.so", "QPPVMPlugin_factory", str(tmp_path / "s.bin"), str(tmp_path / "o.bin"), T, n, False, links)
    per = 3 + n + L.rec_doubles + L.out_doubles
    o = np.fromfile(tmp_path / "o.bin").reshape(T, per)
    status, moved = o[:, 0], o[:, 1]
    eff, rec, out = o[:, 3:3 + n], o[:, 3 + n:3 + n + L.rec_doubles], o[:, 3 + n + L.rec_doubles:]
    assert (status[:-1] == 0).all() and status[-1] != 0
    # failure convention (ref:src/QPPVMPlugin.cpp:246-256, 318-328): tau_qp = 0, command = h, still moves
    assert (moved == 2 * 0 + 1).all() and "SOLVER ERROR!" in stdout
    np.testing.assert_array_equal(eff[-1], h[-1])
    exp = np.zeros((T, L.rec_doubles))
    for t in range(T):
        for ti, b in enumerate(bodies[::-1]):                   # stack order ee_right + ee_left (:177)
            lk = dyn["links"][b]
            J = lk["J"][t][:, 6:]
            e = np.concatenate([dyn["links"][b]["p"][0] - lk["p"][t], _ori_err(dyn["links"][b]["R"][0], lk["R"][t])])
            exp[t, L.off_jc + ti * 6 * n:L.off_jc + (ti + 1) * 6 * n] = J.ravel()
            exp[t, L.off_fee + 6 * ti:L.off_fee + 6 * ti + 6] = 700.0 * e - 70.0 * (J @ qd[t])
        exp[t, L.off_M:L.off_M + n * (n + 1) // 2] = gen.pack_lower(M[t]); exp[t, L.off_h:L.off_h + n] = h[t]
        exp[t, L.off_tauj:L.off_tauj + n] = 5.0 * (q[0] - q[t]) - 2.0 * qd[t]       # _joint_task->setReference(_q) at on_start
        exp[t, L.off_taulim:L.off_taulim + n] = -tmax; exp[t, L.off_taulim + n:L.off_taulim + 2 * n] = tmax
    np.testing.assert_allclose(rec[:-1], exp[:-1], rtol=1e-12, atol=1e-12)
    oo = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, rec[:-1])[0])
    assert (oo["status"] == 0).all()
    assert rel_inf(out[:-1, :n], oo["x"]).max() <= PRIMAL_TOL
    assert rel_inf(eff[:-1], oo["tau"]).max() <= PRIMAL_TOL         # tau_d = tau_qp + h


@pytest.mark.gpu
def test_qppvm_plugin_elbow_and_joint_limit_stack(built, oracle_mod, tmp_path):
    """QPPVM_PLUGIN_STACK=elbows,joint_limits: the drop-in stacks what the reference only constructs
    (ref:src/QPPVMPlugin.cpp:154-171, commented stack :177-178): record packing vs numpy, command vs the oracle."""
    rob = gen.robot_for(29)
    n, T = 29, 6
    desc = Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=FLAG_ELBOW_TASKS | FLAG_JOINT_LIMITS, eps_regularisation=1.0)
    L = layout(desc)
    links = ["arm1_7", "arm2_7", "arm1_4", "arm2_4"]
    bodies = [rob.hand[0], rob.hand[1], rob.hand[0] - 3, rob.hand[1] - 3]
    q, qd, rpy, R0, p0, v0, w0 = _states(rob, T, 11, False)
    dyn = rob.dynamics(q, qd, R0, p0, v0, w0, bodies)
    M = dyn["M"][:, 6:, 6:].copy(); h = dyn["h"][:, 6:].copy()
    tmax = rob.tau_max * 0.6

    def tick(t):
        yield from (q[t], qd[t], rob.q_home, M[t], h[t], tmax, np.zeros(3), np.zeros(3), np.eye(3), np.zeros(3))
        for b in bodies:
            lk = dyn["links"][b]
            yield from (lk["J"][t][:, 6:], lk["Jdqd"][t], lk["R"][t], lk["p"][t], lk["J"][t][:, 6:] @ qd[t])
    _write_states(tmp_path / "s.bin", T, n, tick)
    stdout, _ = _run(built, "libQPPVMPlugin.so", "QPPVMPlugin_factory", str(tmp_path / "s.bin"), str(tmp_path / "o.bin"), T, n, False, links,
                     env={"QPPVM_PLUGIN_STACK": "elbows,joint_limits"})
    per = 3 + n + L.rec_doubles + L.out_doubles
    o = np.fromfile(tmp_path / "o.bin").reshape(T, per)
    eff, rec, out = o[:, 3:3 + n], o[:, 3 + n:3 + n + L.rec_doubles], o[:, 3 + n + L.rec_doubles:]
    assert (o[:, 0] == 0).all()
    # joint limits: model limits home -+ 0.35 shrunk by 10 % of the range on both sides (:120-123), gains k0 * 10, d0 * 20 (:170)
    # with the fake robot's k0 = 1600, d0 = 40; elbows: K = 100, D = 1, reference = pose at construction (tick 0)
    qmin, qmax = rob.q_home - 0.35 + 0.07, rob.q_home + 0.35 - 0.07
    for t in range(T):
        np.testing.assert_allclose(rec[t, L.off_jlim:L.off_jlim + n], 16000.0 * (qmin - q[t]) - 800.0 * qd[t], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(rec[t, L.off_jlim + n:L.off_jlim + 2 * n], 16000.0 * (qmax - q[t]) - 800.0 * qd[t], rtol=1e-12, atol=1e-9)
        for ti, b in enumerate(bodies[2:]):                     # elbow_left + elbow_right (:178)
            lk = dyn["links"][b]
            J = lk["J"][t][:, 6:]
            e = np.concatenate([lk["p"][0] - lk["p"][t], _ori_err(lk["R"][0], lk["R"][t])])
            np.testing.assert_allclose(rec[t, L.off_jelbow + ti * 6 * n:L.off_jelbow + (ti + 1) * 6 * n], J.ravel(), rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(rec[t, L.off_felbow + 6 * ti:L.off_felbow + 6 * ti + 6], 100.0 * e - 1.0 * (J @ qd[t]), rtol=1e-12, atol=1e-10)
    oo = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, rec)[0])
    assert (oo["status"] == 0).all()
    ok = oo["kkt"].max(axis=1) <= 1e-6
    assert ok.sum() >= T - 1
    # the directions neither a hand nor an elbow row sees are regulariser-defined (eps = 2.2e-13): the commanded torques
    # agree where the problem pins them -- the task values of both levels -- and within the conditioning bound elsewhere
    from tests.assemble_np import level_matrices
    for t in np.nonzero(ok)[0]:
        A0 = level_matrices(desc, rec[t], 0)[0]
        A1 = level_matrices(desc, rec[t], 1, out[t, :n])[0]
        for A in (A0, A1):
            np.testing.assert_allclose(A @ out[t, :n], A @ oo["x"][t], rtol=0, atol=1e-6 * max(1.0, np.abs(A @ oo["x"][t]).max()))
    assert rel_inf(out[ok, :n], oo["x"][ok]).max() <= 0.2          # (KKT residual / eps of order one at eps = 2.2e-13)
    np.testing.assert_allclose(eff, out[:, :n] + h, rtol=0, atol=1e-12)        # tau_d = tau_qp + h (:256)
