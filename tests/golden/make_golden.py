"""Generates tests/golden/*.npz: seeded records plus the oracle's solutions (both levels, multipliers,
active masks).  There are NO upstream golden vectors for this path (the reference has no tests and its
solver stack is not installable here: SURVEY.md 4, 8(c)) -- these fixtures pin the oracle against
regressions and travel to the GPU box; they are not reference outputs.  Every stored solution was
verified with the independent numpy KKT certificate (tests/qp_ref.py) when it was generated.

    python tests/golden/make_golden.py [name ...]      (no name: every case)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qppvm_b200 import gen  # noqa: E402
from qppvm_b200.layout import CONFIGS, Desc, KIND_TORQUE, FLAG_COM_TASK, FLAG_JOINT_LIMITS, layout  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.assemble_np import level_matrices  # noqa: E402
from tests.qp_ref import kkt_numpy  # noqa: E402

CASES = {
    "cfg1_forceacc_29dof_2c": (CONFIGS[1]["desc"], gen.config_seed(1)),
    "cfg0_qppvm_29dof_2c_cones_taulim": (CONFIGS[0]["desc"], gen.config_seed(0)),
    "cfg2_qppvm_33dof_4c_cones_taulim": (CONFIGS[2]["desc"], gen.config_seed(2)),
    "torque_29dof_fixed_base": (Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=0, eps_regularisation=1.0), 777),
    # task library: CoM force task at level 1 with cones + torque limits; torque-domain joint limits
    "forceacc_29dof_2c_com_cones_taulim": (Desc(n_a=29, n_contacts=2, flags=FLAG_COM_TASK | 3), 781),
    "torque_29dof_joint_limits": (Desc(kind=KIND_TORQUE, n_a=29, n_contacts=2, flags=FLAG_JOINT_LIMITS, eps_regularisation=1.0), 779),
}
N = 12

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (desc, seed) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        L = layout(desc)
        recs = gen.generate(desc, N, seed)
        out, dg = oracle.solve_batch(desc, recs, mode=oracle.FACTOR_QR, diag=True)
        o = oracle.split_out(desc, out)
        n, nr = L.n_x, L.n_rows
        worst = 0.0
        for i in range(N):
            x0 = dg[i, :n]
            for lev, x, y in ((0, x0, dg[i, n:n + nr]), (1, o["x"][i], dg[i, n + nr:n + 2 * nr])):
                A, b, C, lA, uA, eps = level_matrices(desc, recs[i], lev, x0)
                H, g = A.T @ A + eps * np.eye(n), -A.T @ b
                worst = max(worst, *kkt_numpy(H, g - (eps * x if eps > 0 else 0), C, lA, uA, x, y[:len(lA)]))
        assert (o["status"] == 0).all() and worst < 1e-9, (name, worst)
        np.savez_compressed(os.path.join(here, name + ".npz"), records=recs, x=o["x"], tau=o["tau"],
                            status=o["status"], active=o["active"], x0=dg[:, :n], y0=dg[:, n:n + nr],
                            y1=dg[:, n + nr:n + 2 * nr], eopt=dg[:, n + 2 * nr:],
                            desc=np.array([desc.kind, desc.n_a, desc.n_contacts, desc.flags]),
                            eps_regularisation=desc.eps_regularisation, seed=seed)
        print(name, "kkt(numpy) max", worst, "active rows/problem", np.unpackbits(o["active"].view(np.uint8), axis=1).sum(axis=1).mean())
