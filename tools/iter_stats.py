"""Development aid (run under gpurun): working-set statistics of the solve kernel on configs[2]/[3] and the records
of a large batch that do not end with status OK / KKT <= 1e-6 (saved to gpurun_out/ for a CPU post-mortem)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qppvm_b200 import api, gen, layout  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
START = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = layout.CONFIGS[3]["desc"]
L = layout.layout(d)
s = api.Solver(d)
rob = gen.robot_for(d.n_a)
s.set_robot(rob, (rob.foot + rob.hand)[:d.n_contacts])
bad_recs, bad_idx = [], []
it0, it1, nact = [], [], []
CH = 32768
for c0 in range(START, START + B, CH):
    states = torch.from_numpy(gen.generate_states(d, CH, gen.config_seed(3), c0)).cuda()
    recs = s.records_from_states(states)
    out, _ = s.solve_batch(recs)
    torch.cuda.synchronize()
    g = api.split_out(L, out.cpu().numpy())
    it0.append(g["iters0"]); it1.append(g["iters1"])
    a = g["active"].astype(np.uint32)
    nact.append(sum(((a >> b) & 1).sum(axis=1) for b in range(32)))
    bad = np.nonzero((g["status"] != 0) | ~(g["kkt"].max(axis=1) <= 1e-6))[0]
    if len(bad):
        rc = recs.cpu().numpy()
        for i in bad:
            bad_recs.append(rc[i]); bad_idx.append((c0 + i, int(g["status"][i]), int(g["iters0"][i]), int(g["iters1"][i]), float(g["kkt"][i][0]), float(g["kkt"][i][1])))
it0 = np.concatenate(it0); it1 = np.concatenate(it1); nact = np.concatenate(nact)
print("records", B, "iters0 mean %.2f p99 %d max %d | iters1 mean %.2f p99 %d max %d | active rows (level 1, eq incl.) mean %.2f max %d"
      % (it0.mean(), np.percentile(it0, 99), it0.max(), it1.mean(), np.percentile(it1, 99), it1.max(), nact.mean(), nact.max()))
print("working-set changes beyond the adopted equalities: level 0 %.2f, level 1 %.2f" % (it0.mean() - 6, it1.mean() - 12))
print("hist iters0 (bins of 5)", np.bincount(np.minimum(it0, 60) // 5).tolist())
print("hist iters1 (bins of 5)", np.bincount(np.minimum(it1, 100) // 5).tolist())
print("bad", len(bad_idx), bad_idx[:20])
if bad_recs:
    np.savez("gpurun_out/bad_records_cfg3.npz", recs=np.array(bad_recs), info=np.array(bad_idx))
