"""T3: the CUDA path (through the C-ABI) against the oracle on the same seeded records, for every
BASELINE.json config shape; plus the edge cases of the batch interface and size-independent
properties at the full BASELINE sizes.  Tolerances are north_star's: primal <= 1e-6 relative,
KKT <= 1e-6 at both levels, identical active set at convergence."""
import dataclasses

import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, Desc, KIND_TORQUE, FLAG_COM_TASK, FLAG_ELBOW_TASKS, FLAG_JOINT_LIMITS, layout
from tests.helpers import PRIMAL_TOL, KKT_TOL, compare, rel_inf, mask_differences_are_degenerate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _solve_gpu(torch, desc, recs, diag=True):
    from qppvm_b200 import api
    L = layout(desc)
    s = api.Solver(desc)
    out, dg = s.solve_batch(torch.from_numpy(recs).cuda(), diag=diag)
    torch.cuda.synchronize()
    assert s.kernel_launches >= 1
    return api.split_out(L, out.cpu().numpy()), (api.split_diag(L, dg.cpu().numpy()) if diag else None)


@pytest.mark.parametrize("ci,batch", [(1, 1024), (0, 768), (2, 512)])
def test_parity_vs_oracle(torch_mod, oracle_mod, ci, batch):
    from qppvm_b200 import api
    desc = CONFIGS[ci]["desc"]
    L = layout(desc)
    recs = gen.generate(desc, batch, gen.config_seed(ci))
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == batch
    assert r["primal"] <= PRIMAL_TOL and r["tau"] <= PRIMAL_TOL
    assert r["kkt_gpu"] <= KKT_TOL and r["kkt_oracle"] <= KKT_TOL
    assert r["mask_equal"] == 1.0 and r["strong_active_equal"] == 1.0 and r["strong_sign_equal"] == 1.0
    assert r["eopt"] <= PRIMAL_TOL
    assert rel_inf(gdg["x0"], odg["x0"]).max() <= 1e-4      # level-0 point: eps-defined directions


def test_parity_literal_forceacc_shape(torch_mod, oracle_mod):
    """The reference's own ForceAcc stack: four foot contacts, n_v = 45, no cones / torque limits
    (ref:src/ForceAcc.cpp:58-137); one line in csrc/shapes.def."""
    from qppvm_b200 import api
    desc = Desc(n_a=39, n_contacts=4, flags=0)
    L = layout(desc)
    assert (1, 39, 4, 0) in api.supported_shapes()
    recs = gen.generate(desc, 512, 3939)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == 512
    assert r["primal"] <= PRIMAL_TOL and r["tau"] <= PRIMAL_TOL and r["kkt_gpu"] <= KKT_TOL
    assert r["mask_equal"] == 1.0 and r["strong_active_equal"] == 1.0 and r["strong_sign_equal"] == 1.0


@pytest.mark.parametrize("ci", (1, 2))
def test_parity_with_upstream_semantics_parameters(torch_mod, oracle_mod, ci):
    """The details that varied across OpenSoT versions are explicit parameters of qppvm_desc:
    lambda_solver (g = -lambda A^T W b), one weight per task (W = w I), Postural with or without the six base rows."""
    import dataclasses
    from qppvm_b200 import api
    base = CONFIGS[ci]["desc"]
    desc = dataclasses.replace(base, lambda_solver=0.7, task_weight=(2.0, 0.5, 3.0), postural_actuated_only=1)
    L = layout(desc)
    recs = gen.generate(desc, 384, gen.config_seed(ci) + 5)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == 384
    assert r["primal"] <= PRIMAL_TOL and r["tau"] <= PRIMAL_TOL and r["kkt_gpu"] <= KKT_TOL and r["eopt"] <= PRIMAL_TOL
    assert r["strong_active_equal"] == 1.0 and r["strong_sign_equal"] == 1.0
    # and they do change the answer
    gb, _ = _solve_gpu(torch_mod, base, recs, diag=False)
    assert rel_inf(g["x"], gb["x"]).max() > 1e-3


@pytest.mark.parametrize("n_a,c,flags", [(29, 2, 4), (29, 2, 7)])
def test_parity_full_wrench_variables(torch_mod, oracle_mod, n_a, c, flags):
    """Six variables per contact ("put 6 for full wrench", ref:src/ForceAcc.cpp:67) with the reference's literal wrench
    bounds lb = (-1000, -1000, 10, -1, -1, -1), ub = (1000, 1000, 1000, 1, 1, 1) (ref:src/ForceAcc.cpp:75-76): all six
    rows of every GenericConstraint are real rows, the contact torques enter the dynamics and the torque recovery."""
    from qppvm_b200 import api
    desc = Desc(n_a=n_a, n_contacts=c, flags=flags)
    L = layout(desc)
    assert L.n_x == n_a + 6 + 6 * c
    recs = gen.generate(desc, 384, 6006)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == 384
    assert r["kkt_gpu"] <= KKT_TOL and r["kkt_oracle"] <= KKT_TOL and r["eopt"] <= PRIMAL_TOL
    # Accelerations and contact forces to north_star's 1e-6.  The contact TORQUES carry no task cost: where the
    # dynamics only see a combination of them, the split between contacts is defined by the 2.2e-9 regularisation
    # alone, i.e. to (KKT residual ~1e-12) / 2.2e-9 -- two KKT-exact solutions (same objective to 12 digits, same active
    # set) differ by ~5e-4 there.  Those variables, and the joint torques they feed, are compared to 1e-4.
    sel = np.ones(L.n_x, dtype=bool)
    sel[L.n_v:] = np.tile(np.array([1, 1, 1, 0, 0, 0], dtype=bool), c)
    assert rel_inf(g["x"][:, sel], o["x"][:, sel]).max() <= PRIMAL_TOL
    assert r["primal"] <= 1e-4 and r["tau"] <= 1e-4
    ndiff, tight = mask_differences_are_degenerate(desc, L, recs, g["x"], gdg["x0"], g["active"], o["active"])
    assert tight and ndiff <= 0.02 * 384
    assert r["strong_active_equal"] >= 0.98
    # the contact torques are variables now and sit inside (mostly on) their +-1 bounds
    tq = g["x"][:, L.n_v:].reshape(384, c, 6)[:, :, 3:]
    assert np.abs(tq).max() <= 1.0 + 1e-9 and (np.abs(tq) > 1e-3).any()


@pytest.mark.parametrize("n_a", (29, 39))
def test_parity_torque_kind(torch_mod, oracle_mod, n_a):
    """The literal QPPVMPlugin stack (x = tau, fixed base): ref:src/QPPVMPlugin.cpp:112-188, 201-259."""
    from qppvm_b200 import api
    desc = Desc(kind=KIND_TORQUE, n_a=n_a, n_contacts=2, flags=0, eps_regularisation=1.0)
    L = layout(desc)
    recs = gen.generate(desc, 384, 4242)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == 384
    assert r["primal"] <= PRIMAL_TOL and r["tau"] <= PRIMAL_TOL          # tau = tau_qp + h
    assert r["kkt_gpu"] <= KKT_TOL and r["eopt"] <= PRIMAL_TOL
    # level 1 sits on a degenerate vertex whenever level 0 is bound-limited (its optimality rows pin the
    # level-0 optimum onto the active bounds): the primal point is unique, the multipliers are not.  Active sets
    # must agree except on rows that are tight at the solution.
    ndiff, tight = mask_differences_are_degenerate(desc, L, recs, g["x"], gdg["x0"], g["active"], o["active"])
    assert tight and ndiff <= 0.05 * 384
    # failure convention: tau_qp = 0 -> command = h (QPPVMPlugin.cpp:246-256)
    bad = recs[:4].copy()
    bad[:, L.off_taulim:L.off_taulim + n_a] = 5.0; bad[:, L.off_taulim + n_a:L.off_taulim + 2 * n_a] = -5.0
    gb, _ = _solve_gpu(torch_mod, desc, bad, diag=False)
    assert (gb["status"] != 0).all() and (gb["x"] == 0).all()
    np.testing.assert_array_equal(gb["tau"], bad[:, L.off_h:L.off_h + n_a])


@pytest.mark.parametrize("flags,eps", [(FLAG_JOINT_LIMITS, 1.0), (FLAG_ELBOW_TASKS, 1.0e2), (FLAG_JOINT_LIMITS | FLAG_ELBOW_TASKS, 1.0e2),
                                       (FLAG_ELBOW_TASKS, 1.0)])
def test_parity_torque_task_library(torch_mod, oracle_mod, flags, eps):
    """SURVEY 8(f) row 4, Torque kind: torque-domain JointLimits intersected with the shifted torque limits
    (ref:src/QPPVMPlugin.cpp:169-171) and the elbow tasks as level 1 (ref:src/QPPVMPlugin.cpp:154-166, stack of :177-178)."""
    from qppvm_b200 import api
    desc = Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=flags, eps_regularisation=eps)
    L = layout(desc)
    assert (0, 29, 2, flags) in api.supported_shapes()
    B = 384
    recs = gen.generate(desc, B, 4242)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    assert (g["status"] == o["status"]).all() and (o["status"] == 0).all()
    # With the reference's eps factor (1.0 -> 2.2e-13) the 17 directions that neither a hand nor an elbow row sees are
    # defined by the regulariser alone: a few records per thousand end above the KKT tolerance in BOTH solvers (they are
    # reported as such in the trailer and not counted by the bench); parity is asserted on the records both certify.
    good = (o["kkt"].max(axis=1) <= KKT_TOL) & (g["kkt"].max(axis=1) <= KKT_TOL)
    assert good.mean() >= (0.98 if eps == 1.0 and flags & FLAG_ELBOW_TASKS else 1.0)
    assert rel_inf(gdg["eopt"][good], odg["eopt"][good]).max() <= PRIMAL_TOL
    # the level-0 hand tasks keep their value at level 1 (optimality rows), the bounds hold
    lo = recs[:, L.off_taulim:L.off_taulim + 29] - recs[:, L.off_h:L.off_h + 29]
    hi = recs[:, L.off_taulim + 29:L.off_taulim + 58] - recs[:, L.off_h:L.off_h + 29]
    if flags & FLAG_JOINT_LIMITS:
        lo = np.maximum(lo, recs[:, L.off_jlim:L.off_jlim + 29]); hi = np.minimum(hi, recs[:, L.off_jlim + 29:L.off_jlim + 58])
        jl_binds = (recs[:, L.off_jlim + 29:L.off_jlim + 58] < hi + 1e-12) & (np.abs(g["x"] - hi) < 1e-6)
        assert jl_binds[good].any()                          # a joint-limit bound (not a torque limit) is active somewhere
    tol = KKT_TOL * np.maximum(1.0, np.abs(g["x"][good]).max(axis=1, keepdims=True))     # r_prim is scaled by ||x||_inf
    assert (g["x"][good] >= lo[good] - tol).all() and (g["x"][good] <= hi[good] + tol).all()
    np.testing.assert_allclose(g["tau"][good], g["x"][good] + recs[good, L.off_h:L.off_h + 29], rtol=0, atol=1e-12)
    if flags & FLAG_ELBOW_TASKS:
        # 17 of the 29 directions are only pinned by the regulariser: compare what the problem defines (the task values
        # of both levels) and bound the rest (conditioning limit)
        from tests.assemble_np import level_matrices
        for i in np.nonzero(good)[0][:48]:
            A1, b1, *_ = level_matrices(desc, recs[i], 1, gdg["x0"][i])
            np.testing.assert_allclose(A1 @ g["x"][i], A1 @ o["x"][i], rtol=0, atol=1e-6 * max(1.0, np.abs(b1).max()))
        # measured: 3e-4 at eps factor 1e2 (eps = 2.2e-11), 5e-2 at the reference's 1.0 (eps = 2.2e-13, i.e. KKT residual / eps
        # of order one: neither solver -- nor qpOASES with its 2.2e-7 termination tolerance -- resolves those directions)
        assert rel_inf(g["x"][good], o["x"][good]).max() <= (0.2 if eps == 1.0 else 2e-3)
        nx_g, nx_o = np.linalg.norm(g["x"][good], axis=1), np.linalg.norm(o["x"][good], axis=1)
        assert (np.abs(nx_g - nx_o) <= (1e-2 if eps == 1.0 else 1e-4) * nx_o).all()      # the regulariser's own objective
    else:
        assert rel_inf(g["x"][good], o["x"][good]).max() <= PRIMAL_TOL
        ndiff, tight = mask_differences_are_degenerate(desc, L, recs, g["x"], gdg["x0"], g["active"], o["active"])
        assert tight and ndiff <= 0.10 * B      # (more active bounds than the plain stack: more degenerate vertices)


@pytest.mark.parametrize("flags", (FLAG_COM_TASK, FLAG_COM_TASK | 3))
def test_parity_com_force_task(torch_mod, oracle_mod, flags):
    """the centroidal force task the reference constructs (OpenSoT tasks::force::CoM,
    ref:src/ForceAcc.cpp:103) stacked at level 1.  Its rows sit on the wrench variables: the general dense path of the
    kernel (whitening over all n_x columns, every inequality row through the triangular products)."""
    from qppvm_b200 import api
    desc = Desc(n_a=29, n_contacts=2, flags=flags)
    L = layout(desc)
    assert (1, 29, 2, flags) in api.supported_shapes()
    B = 384
    recs = gen.generate(desc, B, 777)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == B
    assert r["primal"] <= PRIMAL_TOL and r["tau"] <= PRIMAL_TOL
    assert r["kkt_gpu"] <= KKT_TOL and r["kkt_oracle"] <= KKT_TOL and r["eopt"] <= PRIMAL_TOL
    assert r["strong_active_equal"] == 1.0 and r["strong_sign_equal"] == 1.0
    # the task changes the answer: against the same records without it the contact forces move by hundreds of newtons
    base = Desc(n_a=29, n_contacts=2, flags=flags & ~FLAG_COM_TASK)
    Lb = layout(base)
    gb, _ = _solve_gpu(torch_mod, base, np.ascontiguousarray(np.pad(recs[:, :L.off_com], ((0, 0), (0, Lb.rec_doubles - L.off_com)))), diag=False)
    assert np.abs(gb["x"][:, L.n_v:] - g["x"][:, L.n_v:]).max() > 50.0


@pytest.mark.parametrize("ci,steps", [(1, 0), (1, 2), (2, 0), (2, 2)])
def test_parity_number_of_regularisation_steps(torch_mod, oracle_mod, ci, steps):
    """qpOASES' numRegularisationSteps (0 - 2 across OpenSoT versions): every proximal re-solve
    `g <- g - eps x_prev` of kernel and oracle lands on the same point."""
    from qppvm_b200 import api
    desc = dataclasses.replace(CONFIGS[ci]["desc"], n_reg_steps=steps)
    L = layout(desc)
    recs = gen.generate(desc, 256, gen.config_seed(ci) + 5)
    oo, od = oracle_mod.solve_batch(desc, recs, diag=True)
    o, odg = oracle_mod.split_out(desc, oo), api.split_diag(L, od)
    g, gdg = _solve_gpu(torch_mod, desc, recs)
    r = compare(L, g, o, gdg, odg)
    assert r["status_equal"] and r["both_ok"] == 256
    assert r["primal"] <= PRIMAL_TOL and r["tau"] <= PRIMAL_TOL and r["kkt_gpu"] <= KKT_TOL and r["eopt"] <= PRIMAL_TOL
    if steps < 2:
        assert r["mask_equal"] == 1.0 and r["strong_active_equal"] == 1.0
    else:
        # after two proximal steps a few per cent of the records carry a multiplier that is zero to rounding (the point
        # converges onto the unregularised optimum, where those rows are only weakly active): the two solvers may report
        # such a row differently, but only rows that are tight at the solution
        ndiff, tight = mask_differences_are_degenerate(desc, L, recs, g["x"], gdg["x0"], g["active"], o["active"])
        assert tight and ndiff <= 0.06 * 256
    if steps == 2:      # the steps do move the point: against a single step the eps-defined directions shift
        g1, _ = _solve_gpu(torch_mod, dataclasses.replace(desc, n_reg_steps=1), recs, diag=False)
        assert np.abs(g1["x"] - g["x"]).max() > 0.0


def test_infeasible_state_of_the_sharded_workload(torch_mod, oracle_mod):
    """State 696 838 of configs[3] is infeasible (LP-certified in tests/test_oracle_crosscheck.py): kernel and oracle
    both say so, neighbours in the same launch are unaffected, nothing is commanded for it (ref:src/ForceAcc.cpp:189-193)."""
    desc = CONFIGS[3]["desc"]
    L = layout(desc)
    recs = gen.generate(desc, 5, gen.config_seed(3), start=696836)
    o = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, recs)[0])
    g, _ = _solve_gpu(torch_mod, desc, recs, diag=False)
    np.testing.assert_array_equal(g["status"], o["status"])
    assert g["status"].tolist() == [0, 0, 2, 0, 0]
    assert (g["x"][2] == 0).all() and (g["tau"][2] == 0).all() and not np.isfinite(g["kkt"][2]).any()
    ok = g["status"] == 0
    assert rel_inf(g["x"][ok], o["x"][ok]).max() <= PRIMAL_TOL and g["kkt"][ok].max() <= KKT_TOL


def test_empty_single_and_ragged_batches(torch_mod, oracle_mod):
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[1]["desc"]
    L = layout(desc)
    s = api.Solver(desc)
    recs = gen.generate(desc, 301, 99)
    ref = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, recs)[0])
    empty = torch.empty((0, L.rec_doubles), dtype=torch.float64, device="cuda")
    out, _ = s.solve_batch(empty)
    assert out.shape == (0, L.out_doubles)
    for b in (1, 31, 33, 149, 301):                        # below / above a warp, an SM count, ragged
        out, _ = s.solve_batch(torch.from_numpy(recs[:b]).cuda())
        torch.cuda.synchronize()
        g = api.split_out(L, out.cpu().numpy())
        assert (g["status"] == 0).all() and rel_inf(g["x"], ref["x"][:b]).max() <= PRIMAL_TOL


def test_host_and_single_tick_entry_points_match_device_path(torch_mod):
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[0]["desc"]
    L = layout(desc)
    s = api.Solver(desc)
    recs = gen.generate(desc, 5000, 7)                     # > 2 host chunks, ragged tail
    dev, _ = s.solve_batch(torch.from_numpy(recs).cuda())
    torch.cuda.synchronize()
    dev = dev.cpu().numpy()
    host = s.solve_batch_host(recs)
    assert np.array_equal(host, dev)                       # bitwise: problems are independent
    for i in (0, 17, 4999):
        s.reset_warm()                                      # cold start: the same arithmetic as the batch path, bit for bit
        assert np.array_equal(s.solve_one(recs[i]), dev[i])
    # pipelined form: three overlapping calls, one sync
    pin = torch.from_numpy(recs).pin_memory()
    outs = [torch.empty((n, L.out_doubles), dtype=torch.float64).pin_memory() for n in (5000, 1234, 5000)]
    s.solve_batch_host_async_ptr(pin.data_ptr(), outs[0].data_ptr(), 5000)
    s.solve_batch_host_async_ptr(pin[100:].data_ptr(), outs[1].data_ptr(), 1234)
    s.solve_batch_host_async_ptr(pin.data_ptr(), outs[2].data_ptr(), 5000)
    s.host_sync()
    assert np.array_equal(outs[0].numpy(), dev) and np.array_equal(outs[2].numpy(), dev)
    assert np.array_equal(outs[1].numpy(), dev[100:1334])


def test_shard_invariance(torch_mod):
    """T5: any partition of the batch gives bitwise-identical per-problem results."""
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[2]["desc"]
    s = api.Solver(desc)
    recs = torch.from_numpy(gen.generate(desc, 1000, 3)).cuda()
    whole, _ = s.solve_batch(recs)
    parts = [s.solve_batch(recs[a:b].contiguous())[0] for a, b in ((0, 137), (137, 512), (512, 1000))]
    torch.cuda.synchronize()
    assert torch.equal(whole, torch.cat(parts))


@pytest.mark.parametrize("ci", (1, 0, 2))
def test_repeated_launches_are_bitwise_identical(torch_mod, ci):
    """Race proxy (compute-sanitizer is closed on this pool): the kernel has no atomics on data and a fixed
    reduction order, so any run-to-run difference would be a shared-memory race or an uninitialised read."""
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[ci]["desc"]
    s = api.Solver(desc)
    recs = torch.from_numpy(gen.generate(desc, 2500, 77 + ci)).cuda()
    ref, dref = s.solve_batch(recs, diag=True)
    ref, dref = ref.clone(), dref.clone()
    for _ in range(4):
        out, dg = s.solve_batch(recs, diag=True)
        torch.cuda.synchronize()
        assert torch.equal(out, ref) and torch.equal(dg, dref)
    with pytest.raises(api.QPError):                      # records must be 16-byte aligned (TMA bulk copy)
        flat = torch.empty(recs.numel() + 1, dtype=torch.float64, device="cuda")
        mis = flat[1:].view(recs.shape)
        mis.copy_(recs)
        s.solve_batch(mis)


def test_infeasible_and_iteration_limit_status(torch_mod, oracle_mod):
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[1]["desc"]
    L = layout(desc)
    recs = gen.generate(desc, 64, 11)
    bad = recs.copy()
    bad[::2, L.off_fbox + 2] = 50.0; bad[::2, L.off_fbox + 5] = 20.0      # f_z in [50, 20]: empty box
    g, _ = _solve_gpu(torch, desc, bad, diag=False)
    o = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, bad)[0])
    assert (g["status"][::2] == 2).all() and (g["status"][1::2] == 0).all()
    assert np.array_equal(g["status"], o["status"])
    assert (g["x"][::2] == 0).all() and (g["tau"][::2] == 0).all()       # nothing commanded (ForceAcc.cpp:189-193)
    lim = dataclasses.replace(desc, max_iter=7)
    g2, _ = _solve_gpu(torch, lim, recs, diag=False)
    o2 = oracle_mod.split_out(lim, oracle_mod.solve_batch(lim, recs)[0])
    assert (g2["status"] == 1).any() and np.array_equal(g2["status"] == 0, o2["status"] == 0)


def test_non_finite_record_is_flagged_not_propagated(torch_mod):
    from qppvm_b200 import api
    desc = CONFIGS[1]["desc"]
    L = layout(desc)
    recs = gen.generate(desc, 8, 5)
    recs[3, L.off_M + 10] = np.nan
    g, _ = _solve_gpu(torch_mod, desc, recs, diag=False)
    assert g["status"][3] != 0 and (np.delete(g["status"], 3) == 0).all()
    assert np.isfinite(g["x"]).all() and np.isfinite(g["tau"]).all()


@pytest.mark.parametrize("ci", (1, 2))
def test_full_size_properties(torch_mod, oracle_mod, ci):
    """At BASELINE.json's full batch: every solve converges with KKT <= 1e-6 (in-kernel certificate),
    dyn-feas holds, forces respect their box, level 1 keeps the level-0 task value; a random sample is
    compared with the oracle."""
    from qppvm_b200 import api
    desc, batch = CONFIGS[ci]["desc"], CONFIGS[ci]["batch"]
    L = layout(desc)
    recs = gen.generate(desc, batch, gen.config_seed(ci))
    g, gd = _solve_gpu(torch_mod, desc, recs)
    assert (g["status"] == 0).all()
    assert g["kkt"].max() <= KKT_TOL
    nv, c = L.n_v, L.n_c
    f = g["x"][:, nv:].reshape(batch, c, 3)
    fb = recs[:, L.off_fbox:L.off_fbox + 6 * c].reshape(batch, c, 6)
    assert (f >= fb[:, :, :3] - 1e-6).all() and (f <= fb[:, :, 3:] + 1e-6).all()
    M = gen.unpack_lower(recs[:, L.off_M:L.off_M + nv * (nv + 1) // 2], nv)
    Jc = recs[:, L.off_jc:L.off_jc + c * 6 * nv].reshape(batch, c, 6, nv)
    wrench = np.einsum("bckj,bck->bj", Jc[:, :, :3, :], f)
    res = np.einsum("bij,bj->bi", M, g["x"][:, :nv]) + recs[:, L.off_h:L.off_h + nv] - wrench
    assert np.abs(res[:, :6]).max() <= 1e-6 * max(1.0, np.abs(recs[:, L.off_h:L.off_h + nv]).max())
    np.testing.assert_allclose(g["tau"], res[:, 6:], rtol=0, atol=1e-7 * max(1.0, np.abs(res).max()))
    Jw = recs[:, L.off_jwaist:L.off_jwaist + 6 * nv].reshape(batch, 6, nv)
    assert rel_inf(np.einsum("bij,bj->bi", Jw, g["x"][:, :nv]), gd["eopt"]).max() <= 1e-8
    idx = np.random.default_rng(0).choice(batch, 256, replace=False)
    o = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, recs[idx])[0])
    assert rel_inf(g["x"][idx], o["x"]).max() <= PRIMAL_TOL
    assert np.array_equal(g["active"][idx], o["active"])


@pytest.mark.parametrize("ci", (1, 0, 2))
def test_prepared_and_row_by_row_equality_paths_agree(torch_mod, monkeypatch, ci):
    """The solve kernel adopts the equality working set the prepare kernel orthogonalised; a dependent row (or
    non-finite data) makes it fall back to adding the equality rows one by one.  QPPVM_ROWWISE_EQUALITIES=1 at
    create forces that fallback for every problem: both paths must give the same solutions and active sets."""
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[ci]["desc"]
    L = layout(desc)
    recs = torch.from_numpy(gen.generate(desc, 600, 1234 + ci)).cuda()
    fast = api.Solver(desc)
    monkeypatch.setenv("QPPVM_ROWWISE_EQUALITIES", "1")
    slow = api.Solver(desc)
    monkeypatch.delenv("QPPVM_ROWWISE_EQUALITIES")
    a, _ = fast.solve_batch(recs)
    b, _ = slow.solve_batch(recs)
    torch.cuda.synchronize()
    ga, gb = api.split_out(L, a.cpu().numpy()), api.split_out(L, b.cpu().numpy())
    assert (ga["status"] == 0).all() and np.array_equal(ga["status"], gb["status"])
    assert rel_inf(ga["x"], gb["x"]).max() <= 1e-9 and rel_inf(ga["tau"], gb["tau"]).max() <= 1e-9
    assert np.array_equal(ga["active"], gb["active"]) and max(ga["kkt"].max(), gb["kkt"].max()) <= KKT_TOL
    assert not torch.equal(a, b)                           # (the two paths really are different arithmetic)


def test_one_handle_moved_between_streams(torch_mod):
    """The caller's-stream launch slot (work counter, prepare workspace) is reused across calls: calls issued
    back to back on different streams, without any synchronisation in between, must not disturb each other."""
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[1]["desc"]
    s = api.Solver(desc)
    a = torch.from_numpy(gen.generate(desc, 3000, 501)).cuda()
    b = torch.from_numpy(gen.generate(desc, 2000, 502)).cuda()
    ref_a, ref_b = s.solve_batch(a)[0].clone(), s.solve_batch(b)[0].clone()
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(6):
        st, rec = (sa, a) if rep % 2 == 0 else (sb, b)
        outs.append(s.solve_batch(rec, stream=st)[0])
    torch.cuda.synchronize()
    for rep, o in enumerate(outs):
        assert torch.equal(o, ref_a if rep % 2 == 0 else ref_b)
