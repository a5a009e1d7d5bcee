"""Development aid (run under gpurun): GPU kernel vs oracle on every instantiated shape, verbose."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from qppvm_b200 import api, gen, layout  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.helpers import compare  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
TORQUE = layout.Desc(kind=layout.KIND_TORQUE, n_a=29, n_contacts=2, flags=0, eps_regularisation=1.0)
for ci in (1, 0, 2, 7):
    d = TORQUE if ci == 7 else layout.CONFIGS[ci]["desc"]
    L = layout.layout(d)
    recs = gen.generate(d, B, gen.config_seed(ci))
    t = time.time(); oo, od = oracle.solve_batch(d, recs, diag=True); t_or = time.time() - t
    s = api.Solver(d)
    rd = torch.from_numpy(recs).cuda()
    out, dg = s.solve_batch(rd, diag=True)
    torch.cuda.synchronize()
    g = api.split_out(L, out.cpu().numpy()); o = oracle.split_out(d, oo)
    gd = api.split_diag(L, dg.cpu().numpy()); odg = api.split_diag(L, od)
    r = compare(L, g, o, gd, odg)
    print("config", ci, json.dumps(r))
    print("  status gpu", np.bincount(g["status"], minlength=4), "oracle", np.bincount(o["status"], minlength=4))
    print("  iters gpu %.1f %.1f oracle %.1f %.1f" % (g["iters0"].mean(), g["iters1"].mean(), o["iters0"].mean(), o["iters1"].mean()))
    bad = np.nonzero((g["status"] != 0) | (np.abs(g["x"] - o["x"]).max(axis=1) > 1e-6 * np.maximum(1, np.abs(o["x"]).max(axis=1))))[0]
    for i in bad[:5]:
        print("  bad", i, "status", g["status"][i], o["status"][i], "it", g["iters0"][i], g["iters1"][i], "kkt", g["kkt"][i], o["kkt"][i],
              "dx", np.abs(g["x"][i] - o["x"][i]).max(), "dx0", np.abs(gd["x0"][i] - odg["x0"][i]).max(),
              "mask", g["active"][i], o["active"][i], "fail(slack,|bound|,k)", gd["eopt"][i][3:6])
    # timing
    for nb in (4096, 32768):
        big = rd.repeat((nb + B - 1) // B, 1)[:nb].contiguous()
        outb = torch.empty((nb, L.out_doubles), dtype=torch.float64, device="cuda")
        s.solve_batch(big, out=outb); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_batch(big, out=outb); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("  batch %d: %.3f ms  %.0f solves/s  (oracle %d threads: %.0f solves/s)" % (nb, ms, nb / ms * 1e3, oracle.num_threads(), B / t_or))
    print("  fp64 peak TFLOP/s", s.fp64_peak_tflops(), "ctas/SM via launches", s.kernel_launches)
