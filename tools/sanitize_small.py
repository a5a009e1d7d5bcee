"""Tiny workload for compute-sanitizer (T6): a few problems of every instantiated shape through the C-ABI."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qppvm_b200 import api, gen, layout  # noqa: E402

TORQUE = layout.Desc(kind=layout.KIND_TORQUE, n_a=29, n_contacts=2, flags=0, eps_regularisation=1.0)
for name, d in (("cfg1", layout.CONFIGS[1]["desc"]), ("cfg0", layout.CONFIGS[0]["desc"]), ("cfg2", layout.CONFIGS[2]["desc"]), ("torque29", TORQUE)):
    L = layout.layout(d)
    recs = gen.generate(d, 12, 31)
    s = api.Solver(d)
    out, dg = s.solve_batch(torch.from_numpy(recs).cuda(), diag=True)
    torch.cuda.synchronize()
    g = api.split_out(L, out.cpu().numpy())
    one = s.solve_one(recs[0])
    host = s.solve_batch_host(recs)
    print(name, "status", g["status"].tolist(), "kkt", float(g["kkt"].max()), "host==dev", bool(np.array_equal(host, out.cpu().numpy())))
