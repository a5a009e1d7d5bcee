"""Record / output layout of the whole-body QP hot path (host-side mirror of
``include/qppvm_b200.h``; ``tests/test_abi.py`` checks both agree field by field).

Variables and rows follow the reference:
  * ForceAcc kind: x = [qddot (n_v) | f_0 .. f_{c-1} (3 each)]   ref:src/ForceAcc.cpp:63-70
    rows: dyn-feas 6 | wrench box 6c | [cones 5c] | [torque n_a] | level-1 optimality 6
    (stack order ``<< _dyn_feas << wrench_bounds[i]``, ref:src/ForceAcc.cpp:131-133)
  * Torque kind: x = tau (n)                                       ref:src/QPPVMPlugin.cpp:177-179
    rows: simple bounds n | level-1 optimality 6
"""
from __future__ import annotations

from dataclasses import dataclass, field

KIND_TORQUE = 0
KIND_FORCEACC = 1
FLAG_FRICTION_CONES = 1
FLAG_TORQUE_LIMITS = 2
FLAG_FULL_WRENCH = 4      # 6 variables per contact ("put 6 for full wrench", ref:src/ForceAcc.cpp:67)
FLAG_COM_TASK = 32        # ForceAcc kind: the centroidal force task joins level 1 (ref:src/ForceAcc.cpp:103)
FLAG_JOINT_LIMITS = 8     # Torque kind: torque-domain JointLimits bounds (ref:src/QPPVMPlugin.cpp:169-171)
FLAG_ELBOW_TASKS = 16     # Torque kind: level 1 = elbow_left + elbow_right (ref:src/QPPVMPlugin.cpp:154-166, 177-178)
STATUS_OK, STATUS_MAX_ITER, STATUS_INFEASIBLE, STATUS_NUMERIC = 0, 1, 2, 3
QPOASES_EPS = 2.221e-16
QPOASES_EPS_REG = 1.0e3 * QPOASES_EPS
INFTY = 1.0e20
M0 = 6


@dataclass(frozen=True)
class Desc:
    """Mirror of ``qppvm_desc``."""
    kind: int = KIND_FORCEACC
    n_a: int = 29
    n_contacts: int = 2
    flags: int = 0
    eps_regularisation: float = 1.0e4   # ref:src/ForceAcc.cpp:137 (QPPVMPlugin.cpp:188 uses 1.0)
    n_reg_steps: int = 1                # qpOASES setToMPC(): numRegularisationSteps = 1
    max_iter: int = 132                 # nWSR
    device: int = 0
    postural_actuated_only: int = 0     # later OpenSoT versions drop the 6 base rows of Postural
    lambda_solver: float = 1.0          # g = -lambda A^T W b
    task_weight: tuple = (1.0, 1.0, 1.0)   # W = w I per task: waist, postural, contact Cartesian

    @property
    def eps(self) -> float:
        return self.eps_regularisation * QPOASES_EPS_REG


@dataclass(frozen=True)
class Layout:
    n_a: int
    n_v: int
    n_c: int
    n_x: int
    n_rows: int
    row_dyn: int
    row_box: int
    row_cone: int
    row_tau: int
    row_opt: int
    off_jwaist: int
    off_jc: int
    off_M: int
    off_h: int
    off_jdqd: int
    off_rhs: int
    off_taulim: int
    off_cone: int
    off_fbox: int
    off_fee: int
    off_tauj: int
    rec_doubles: int
    out_bytes: int
    diag_doubles: int
    off_jlim: int = -1
    off_jelbow: int = -1
    off_felbow: int = -1
    off_com: int = -1
    FIELDS = ("n_a", "n_v", "n_c", "n_x", "n_rows", "row_dyn", "row_box", "row_cone", "row_tau",
              "row_opt", "off_jwaist", "off_jc", "off_M", "off_h", "off_jdqd", "off_rhs",
              "off_taulim", "off_cone", "off_fbox", "off_fee", "off_tauj", "rec_doubles",
              "out_bytes", "diag_doubles", "off_jlim", "off_jelbow", "off_felbow", "off_com")

    @property
    def out_doubles(self) -> int:
        return self.out_bytes // 8

    def algorithmic_bytes(self) -> int:
        """unpadded record bytes + output bytes per solve."""
        return 8 * self.rec_doubles_unpadded + self.out_bytes

    @property
    def rec_doubles_unpadded(self) -> int:
        return self._unpadded

    _unpadded: int = field(default=0, repr=False)


def layout(desc: Desc) -> Layout:
    if desc.n_a < 1 or desc.n_a > 58:
        raise ValueError("n_a out of range")
    off_jlim = off_jelbow = off_felbow = off_com = -1
    if desc.kind == KIND_FORCEACC:
        c = desc.n_contacts
        if c < 1 or c > 4:
            raise ValueError("n_contacts out of range")
        n_a, n_v = desc.n_a, desc.n_a + 6
        wd = 6 if desc.flags & FLAG_FULL_WRENCH else 3
        n_x = n_v + wd * c
        if n_x > 64:
            raise ValueError("n_x > 64 unsupported")
        cones = bool(desc.flags & FLAG_FRICTION_CONES)
        tl = bool(desc.flags & FLAG_TORQUE_LIMITS)
        row = 0
        row_dyn = row; row += 6
        row_box = row; row += 6 * c
        row_cone = row if cones else -1; row += 5 * c if cones else 0
        row_tau = row if tl else -1; row += n_a if tl else 0
        row_opt = row; row += M0
        off = 0
        off_jwaist = off; off += 6 * n_v
        off_jc = off; off += c * 6 * n_v
        off_M = off; off += n_v * (n_v + 1) // 2
        off_h = off; off += n_v
        off_jdqd = off; off += 6 * (1 + c)
        off_rhs = off; off += 6 * (1 + c) + n_v
        off_taulim = off if tl else -1; off += 2 * n_a if tl else 0
        off_cone = off if cones else -1; off += 10 * c if cones else 0
        off_fbox = off; off += 2 * wd * c
        if desc.flags & FLAG_COM_TASK:
            off_com = off; off += 6 * wd * c + 6
        if desc.flags & ~(FLAG_FRICTION_CONES | FLAG_TORQUE_LIMITS | FLAG_FULL_WRENCH | FLAG_COM_TASK):
            raise ValueError("ForceAcc kind: unknown flag")
        off_fee = off_tauj = -1
    elif desc.kind == KIND_TORQUE:
        if desc.n_contacts != 2 or desc.flags & ~(FLAG_JOINT_LIMITS | FLAG_ELBOW_TASKS):
            raise ValueError("torque kind: n_contacts must be 2, flags a subset of JOINT_LIMITS | ELBOW_TASKS")
        c = 2
        n_a = n_v = n_x = desc.n_a
        row_dyn = row_cone = row_tau = -1
        row_box = 0
        row_opt = n_x
        row = n_x + M0
        off = 0
        off_jwaist = -1
        off_jc = off; off += 2 * 6 * n_v
        off_M = off; off += n_v * (n_v + 1) // 2
        off_h = off; off += n_v
        off_jdqd = off_rhs = -1
        off_fee = off; off += 12
        off_tauj = off; off += n_v
        off_taulim = off; off += 2 * n_v
        off_cone = off_fbox = -1
        if desc.flags & FLAG_JOINT_LIMITS:
            off_jlim = off; off += 2 * n_v
        if desc.flags & FLAG_ELBOW_TASKS:
            off_jelbow = off; off += 12 * n_v
            off_felbow = off; off += 12
    else:
        raise ValueError("unknown kind")
    if row > 128:
        raise ValueError("more than 128 constraint rows")
    unp = off
    rec = off + (off & 1)
    return Layout(n_a=n_a, n_v=n_v, n_c=c, n_x=n_x, n_rows=row, row_dyn=row_dyn, row_box=row_box,
                  row_cone=row_cone, row_tau=row_tau, row_opt=row_opt, off_jwaist=off_jwaist,
                  off_jc=off_jc, off_M=off_M, off_h=off_h, off_jdqd=off_jdqd, off_rhs=off_rhs,
                  off_taulim=off_taulim, off_cone=off_cone, off_fbox=off_fbox, off_fee=off_fee,
                  off_tauj=off_tauj, rec_doubles=rec, out_bytes=8 * (n_x + n_a) + 32,
                  diag_doubles=n_x + 2 * row + M0, off_jlim=off_jlim, off_jelbow=off_jelbow, off_felbow=off_felbow, off_com=off_com,
                  _unpadded=unp)


# The five BASELINE.json configs (SURVEY.md 8(a) "Per-config QP dimensions").
CONFIGS = {
    0: dict(name="single-tick QPPVM, COMAN-like 29-DoF, 2 contacts, cones+tau-limits",
            desc=Desc(n_a=29, n_contacts=2, flags=FLAG_FRICTION_CONES | FLAG_TORQUE_LIMITS), batch=1),
    1: dict(name="ForceAcc batched, 4096 states, COMAN-like 29-DoF, 2 contacts",
            desc=Desc(n_a=29, n_contacts=2, flags=0), batch=4096),
    2: dict(name="QPPVM batched, 65536 states, WALK-MAN-like 33-DoF, 4 contacts, cones+tau-limits",
            desc=Desc(n_a=33, n_contacts=4, flags=FLAG_FRICTION_CONES | FLAG_TORQUE_LIMITS), batch=65536),
    3: dict(name="QPPVM batched, 1M states sharded, WALK-MAN-like 33-DoF, 4 contacts, cones+tau-limits",
            desc=Desc(n_a=33, n_contacts=4, flags=FLAG_FRICTION_CONES | FLAG_TORQUE_LIMITS), batch=1 << 20),
    4: dict(name="latency mode: single QP per tick, COMAN-like 29-DoF, 2 contacts, cones+tau-limits",
            desc=Desc(n_a=29, n_contacts=2, flags=FLAG_FRICTION_CONES | FLAG_TORQUE_LIMITS), batch=1),
}
