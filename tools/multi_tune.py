"""Root-resident multi-GPU leg (qppvm_multi_solve_batch: NCCL scatter -> solve -> gather) over the root's share of the
batch and the pipeline chunk:  python tools/multi_tune.py G "share,chunk share,chunk ..."   (configs[3], 2^20 states)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qppvm_b200 import api, gen
from qppvm_b200.layout import CONFIGS, layout

G = int(sys.argv[1])
settings = [tuple(s.split(",")) for s in sys.argv[2].split()]
desc = CONFIGS[3]["desc"]; L = layout(desc); B = CONFIGS[3]["batch"]
rob = gen.robot_for(desc.n_a); contacts = (rob.foot + rob.hand)[:desc.n_contacts]
s0 = api.Solver(desc); s0.set_robot(rob, contacts)
import numpy as np
cache = "/dev/shm/qppvm_states_cfg3.npy"      # (several invocations with different NCCL_* environments share the states)
if os.path.exists(cache): st = np.load(cache)
else:
    st = gen.generate_states(desc, B, gen.config_seed(3)); np.save(cache, st)
recs = s0.records_from_states(torch.from_numpy(st).cuda()); torch.cuda.synchronize()
out = torch.empty((B, L.out_doubles), dtype=torch.float64, device="cuda:0")
ref = None
for share, chunk in settings:
    os.environ["QPPVM_MULTI_ROOT_SHARE"] = share; os.environ["QPPVM_MULTI_CHUNK"] = chunk
    m = api.MultiSolver(desc, list(range(G))); m.set_robot(rob, contacts)
    for _ in range(2): m.solve_batch(recs, out=out)
    t0 = time.perf_counter()
    for _ in range(4): m.solve_batch(recs, out=out)
    dt = (time.perf_counter() - t0) / 4
    if ref is None: ref = out.clone()
    print(os.environ.get("TUNE_TAG", ""), "G=%d share=%s chunk=%s  %.2f ms/step  %.2f M solves/s  bitwise_equal_to_first=%s" % (G, share, chunk, dt * 1e3, B / dt / 1e6, bool(torch.equal(out, ref))), flush=True)
    dst = torch.from_numpy(st).cuda()
    for _ in range(2): m.solve_states(dst, out=out)
    t0 = time.perf_counter()
    for _ in range(4): m.solve_states(dst, out=out)
    dts = (time.perf_counter() - t0) / 4
    print("      states on the root: %.2f ms/step  %.2f M solves/s  bitwise_equal=%s" % (dts * 1e3, B / dts / 1e6, bool(torch.equal(out, ref))), flush=True)
    del dst
    m.close()
