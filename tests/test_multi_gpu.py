"""T5 on hardware: the batch sharded over every visible GPU from one process through
qppvm_multi_* (block split, NCCL scatter / gather pipelined against the solves) gives bitwise the results of one GPU --
the problems are independent (ref:src/QPPVMPlugin.cpp:246, ref:src/ForceAcc.cpp:189).  On a one-GPU box the same entry
points run with G = 1."""
import numpy as np
import pytest

from qppvm_b200 import gen
from qppvm_b200.layout import CONFIGS, layout

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _records(torch, desc, n):
    """configs[3] states through the device front end (seconds instead of minutes for 10^5+ records)."""
    from qppvm_b200 import api
    s = api.Solver(desc)
    rob = gen.robot_for(desc.n_a)
    s.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
    st = gen.generate_states(desc, n, gen.config_seed(3))
    recs = s.records_from_states(torch.from_numpy(st).cuda())
    torch.cuda.synchronize()
    return s, st, recs


@pytest.mark.parametrize("batch", (40000, 1 << 17))
def test_sharded_solve_is_bitwise_the_single_gpu_solve(torch_mod, monkeypatch, batch):
    from qppvm_b200 import api
    torch = torch_mod
    desc = CONFIGS[3]["desc"]
    L = layout(desc)
    monkeypatch.setenv("QPPVM_MULTI_CHUNK", "4096")        # several pipeline chunks per GPU, ragged last one
    s, st, recs = _records(torch, desc, batch)
    ref, _ = s.solve_batch(recs)
    torch.cuda.synchronize()
    g = api.split_out(L, ref.cpu().numpy())
    assert (g["status"] == 0).all() and g["kkt"].max() <= 1e-6
    n_gpu = torch.cuda.device_count()
    for devices in ([0], list(range(n_gpu))) if n_gpu > 1 else ([0],):
        m = api.MultiSolver(desc, devices)
        out = m.solve_batch(recs)
        assert torch.equal(out, ref), "devices %s" % devices
        if len(devices) > 1:
            assert m.nccl_calls > 0
        # host buffers: every GPU pulls its own block
        pin = recs.cpu().pin_memory()
        hout = torch.empty((batch, L.out_doubles), dtype=torch.float64).pin_memory()
        m.solve_batch_host_ptr(pin.data_ptr(), hout.data_ptr(), batch)
        assert torch.equal(hout, ref.cpu())
        rob = gen.robot_for(desc.n_a)
        m.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
        hst = torch.from_numpy(st).pin_memory()
        m.solve_states_host_ptr(hst.data_ptr(), hout.data_ptr(), batch)
        assert torch.equal(hout, ref.cpu())
        # states on the root GPU: they are what NCCL scatters, every GPU runs the front end on its chunks
        nc0 = m.nccl_calls
        out_s = m.solve_states(torch.from_numpy(st).cuda())
        assert torch.equal(out_s, ref), "states form, devices %s" % devices
        if len(devices) > 1:
            assert m.nccl_calls > nc0
        m.close()


def test_multi_rejects_bad_arguments(torch_mod):
    from qppvm_b200 import api
    desc = CONFIGS[1]["desc"]
    with pytest.raises(api.QPError):
        api.MultiSolver(desc, [0, 0])
    m = api.MultiSolver(CONFIGS[2]["desc"], [0])
    with pytest.raises(api.QPError):                       # the states form needs the robot tables
        m.solve_states(torch_mod.zeros((4, 128), dtype=torch_mod.float64, device="cuda:0"))
    m.close()
    with pytest.raises(api.QPError):
        api.MultiSolver(desc, [torch_mod.cuda.device_count() + 3])
