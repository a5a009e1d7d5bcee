"""Hot start: the reference keeps one QPOases_sot alive across ticks
(ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:64, ref:src/QPPVMPlugin.cpp:246, ref:src/ForceAcc.cpp:189), so every tick
starts from the previous tick's working set.  Here: warm-start masks per problem (batch path) and the resident
latency-mode chain behind qppvm_solve_one.  The minimiser is unique, so every hot-started solve must agree with the
COLD oracle within north_star's tolerances, with the same strongly active set."""
import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, layout
from tests.helpers import PRIMAL_TOL, KKT_TOL, rel_inf, mask_differences_are_degenerate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("ci", (2, 0, 1))
def test_warm_batch_matches_cold_and_saves_iterations(torch_mod, oracle_mod, ci):
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[ci]["desc"]
    L = layout(desc)
    recs = gen.generate(desc, 768, gen.config_seed(ci))
    s = api.Solver(desc)
    d = torch.from_numpy(recs).cuda()
    cold, _ = s.solve_batch(d)
    warm = torch.zeros((768, 8), dtype=torch.int32, device="cuda")
    first = s.solve_batch_warm(d, warm)
    torch.cuda.synchronize()
    assert torch.equal(first, cold)                        # all-zero masks: the cold path, bit for bit
    gw = warm.cpu().numpy().view(np.uint32)
    c = api.split_out(L, cold.cpu().numpy())
    assert np.array_equal(gw[:, 4:], c["active"])          # level-1 working set == the trailer's mask
    assert (gw[:, :4].any(axis=1)).all()                   # level 0 always holds its equality rows
    # slightly different problems (the next control tick), started from those working sets
    recs2 = recs.copy()
    recs2[:, L.off_rhs:L.off_rhs + 6] *= 1.001
    d2 = torch.from_numpy(recs2).cuda()
    hot = s.solve_batch_warm(d2, warm.clone())
    cold2, dg2 = s.solve_batch(d2, diag=True)
    torch.cuda.synchronize()
    h, c2 = api.split_out(L, hot.cpu().numpy()), api.split_out(L, cold2.cpu().numpy())
    o = oracle_mod.split_out(desc, oracle_mod.solve_batch(desc, recs2)[0])
    assert (h["status"] == 0).all() and (o["status"] == 0).all()
    assert rel_inf(h["x"], o["x"]).max() <= PRIMAL_TOL and rel_inf(h["tau"], o["tau"]).max() <= PRIMAL_TOL
    assert h["kkt"].max() <= KKT_TOL
    x0 = api.split_diag(L, dg2.cpu().numpy())["x0"]
    ndiff, tight = mask_differences_are_degenerate(desc, L, recs2, h["x"], x0, h["active"], o["active"])
    assert tight and ndiff <= 0.02 * 768
    it_hot = (h["iters0"] + h["iters1"]).mean()
    it_cold = (c2["iters0"] + c2["iters1"]).mean()
    assert it_hot <= it_cold                               # never more working-set changes than a cold start, on average
    if ci != 1:
        assert it_hot < 0.97 * it_cold


def _tick_sequence(desc, n_ticks, seed):
    """One robot followed over n_ticks control periods: a slow random walk of q and qdot around a sampled state."""
    st0 = gen.generate_states(desc, 1, seed)[0]
    rng = np.random.default_rng(seed)
    na = desc.n_a
    states = np.repeat(st0[None], n_ticks, axis=0)
    states[:, :na] += np.cumsum(rng.normal(0.0, 2e-3, (n_ticks, na)), axis=0)
    states[:, na:2 * na] += np.cumsum(rng.normal(0.0, 5e-3, (n_ticks, na)), axis=0)
    return gen.records_from_states(desc, states)


def test_thousand_tick_sequence_hot_started(torch_mod, oracle_mod):
    """configs[4]: 1 000 consecutive ticks through qppvm_solve_one (resident chain, hot start from the previous tick);
    every tick against the cold oracle."""
    from qppvm_b200 import api
    desc = CONFIGS[4]["desc"]
    L = layout(desc)
    recs = _tick_sequence(desc, 1000, 77)
    o_out, o_dg = oracle_mod.solve_batch(desc, recs, diag=True)
    o = oracle_mod.split_out(desc, o_out)
    assert (o["status"] == 0).all()
    s = api.Solver(desc)
    outs = np.empty((1000, L.out_doubles))
    for i in range(1000):
        s.solve_one(recs[i], outs[i])
    g = api.split_out(L, outs)
    assert (g["status"] == 0).all()
    assert rel_inf(g["x"], o["x"]).max() <= PRIMAL_TOL and rel_inf(g["tau"], o["tau"]).max() <= PRIMAL_TOL
    assert g["kkt"].max() <= KKT_TOL
    x0 = api.split_diag(L, o_dg)["x0"]
    ndiff, tight = mask_differences_are_degenerate(desc, L, recs, g["x"], x0, g["active"], o["active"])
    assert tight and ndiff <= 20
    # cold start of the same ticks: more working-set changes
    cold_it = 0
    for i in range(0, 1000, 10):
        s.reset_warm()
        t = api.split_out(L, s.solve_one(recs[i])[None])
        cold_it += int(t["iters0"][0] + t["iters1"][0])
    hot_it = int((g["iters0"][::10] + g["iters1"][::10]).sum())
    assert hot_it < cold_it


def test_resident_chain_restarts_after_idle_and_matches_launch_path(torch_mod, monkeypatch):
    """The resident servers leave after QPPVM_TICK_IDLE_US without a tick and come back on the next one; the launch-per-
    tick fallback (QPPVM_RESIDENT=0) gives the same bits."""
    import time
    from qppvm_b200 import api
    desc = CONFIGS[4]["desc"]
    recs = gen.generate(desc, 8, 5)
    monkeypatch.setenv("QPPVM_TICK_IDLE_US", "2000")
    s = api.Solver(desc)
    a = [s.solve_one(recs[i]).copy() for i in range(4)]
    time.sleep(0.05)                                       # > idle time: the chain has left
    b = [s.solve_one(recs[i]).copy() for i in range(4, 8)]
    torch_mod.cuda.synchronize()                           # returns: no resident kernel outlives its idle time
    monkeypatch.setenv("QPPVM_RESIDENT", "0")
    s2 = api.Solver(desc)
    ref = [s2.solve_one(recs[i]).copy() for i in range(8)]
    for x, y in zip(a + b, ref):
        assert np.array_equal(x, y)
