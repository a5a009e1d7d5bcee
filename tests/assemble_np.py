"""Independent numpy restatement of the OpenSoT stack assembly used to pin the
oracle's C assembly; written from the reference call sites, shares no code with oracle/ or csrc/."""
from __future__ import annotations

import numpy as np

from qppvm_b200.gen import unpack_lower
from qppvm_b200.layout import Desc, KIND_FORCEACC, FLAG_COM_TASK, FLAG_ELBOW_TASKS, FLAG_JOINT_LIMITS, layout, QPOASES_EPS_REG, INFTY


def level_matrices(desc: Desc, rec: np.ndarray, level: int, x0=None):
    """(A, b, C, lA, uA, eps) of priority level `level` for one record."""
    L = layout(desc)
    n, nv, c, na = L.n_x, L.n_v, L.n_c, L.n_a
    if desc.kind == KIND_FORCEACC:
        Jw = rec[L.off_jwaist:L.off_jwaist + 6 * nv].reshape(6, nv)
        Jc = rec[L.off_jc:L.off_jc + c * 6 * nv].reshape(c, 6, nv)
        M = unpack_lower(rec[L.off_M:L.off_M + nv * (nv + 1) // 2], nv)
        h = rec[L.off_h:L.off_h + nv]
        jdqd = rec[L.off_jdqd:L.off_jdqd + 6 * (1 + c)]
        rhs = rec[L.off_rhs:L.off_rhs + 6 * (1 + c) + nv]
        Z = np.zeros
        wd = (n - nv) // c                       # 3 force components per contact, 6 with full wrenches (ForceAcc.cpp:67)
        # cost of a level: sum_k w_k/2 ||A_k x - lambda b_k||^2  (H = A^T W A, g = -lambda A^T W b: SURVEY App. A.2)
        lam = desc.lambda_solver
        w_waist, w_post, w_cont = (np.sqrt(w) for w in desc.task_weight)
        if level == 0:   # _waist_task (ForceAcc.cpp:118-122)
            A = w_waist * np.hstack([Jw, Z((6, wd * c))]); b = w_waist * lam * (rhs[:6] - jdqd[:6])
        else:            # _postural_task + feet_cart_aggr (ForceAcc.cpp:131)
            P = np.eye(nv) * w_post
            if desc.postural_actuated_only:   # later OpenSoT versions: no postural rows for the floating base (A.6)
                P[:6] = 0.0
            A = np.vstack([np.hstack([P, Z((nv, wd * c))])] +
                          [w_cont * np.hstack([Jc[i], Z((6, wd * c))]) for i in range(c)])
            b = np.concatenate([P @ (lam * rhs[6 * (1 + c):])] +
                               [w_cont * lam * (rhs[6 * (1 + i):6 * (2 + i)] - jdqd[6 * (1 + i):6 * (2 + i)]) for i in range(c)])
            if desc.flags & FLAG_COM_TASK:   # ... + _com_task (constructed at ForceAcc.cpp:103): rows on the wrench variables
                blk = rec[L.off_com:L.off_com + 6 * wd * c + 6]
                A = np.vstack([A, np.hstack([Z((6, nv)), blk[:6 * wd * c].reshape(6, wd * c)])])
                b = np.concatenate([b, lam * blk[6 * wd * c:]])
        rows, lo, hi = [], [], []
        # DynamicFeasibility: (M qdd + h - sum J_i^T [f_i; 0])[0:6] = 0
        D = np.hstack([M[:6]] + [-Jc[i][:wd, :6].T for i in range(c)])
        rows.append(D); lo.append(-h[:6]); hi.append(-h[:6])
        for i in range(c):   # wrench_i = force_i / Zero(3)  in [lb, ub]  (ForceAcc.cpp:74-76, 81, 91-95)
            W = Z((6, n)); W[:wd, nv + wd * i:nv + wd * i + wd] = np.eye(wd)
            fb = rec[L.off_fbox + 2 * wd * i:L.off_fbox + 2 * wd * i + 2 * wd]
            rows.append(W); lo.append(np.concatenate([fb[:wd], -np.ones(6 - wd)])); hi.append(np.concatenate([fb[wd:], np.ones(6 - wd)]))
        if L.row_cone >= 0:
            for i in range(c):
                blk = rec[L.off_cone + 10 * i:L.off_cone + 10 * i + 10]
                R, mu = blk[:9].reshape(3, 3), blk[9] / np.sqrt(2.0)
                Ci = np.array([[1, 0, -mu], [-1, 0, -mu], [0, 1, -mu], [0, -1, -mu], [0, 0, -1.0]])
                F = Z((5, n)); F[:, nv + wd * i:nv + wd * i + 3] = Ci @ R.T
                rows.append(F); lo.append(np.full(5, -INFTY)); hi.append(np.zeros(5))
        if L.row_tau >= 0:
            T = np.hstack([M[6:]] + [-Jc[i][:wd, 6:].T for i in range(c)])
            tl = rec[L.off_taulim:L.off_taulim + 2 * na]
            rows.append(T); lo.append(tl[:na] - h[6:]); hi.append(tl[na:] - h[6:])
        if level == 1:   # optimality rows: the level-0 task rows themselves (a task weight does not change the set they define)
            A0 = np.hstack([Jw, Z((6, wd * c))])
            rows.append(A0); lo.append(A0 @ x0); hi.append(A0 @ x0)
        eps = desc.eps_regularisation * QPOASES_EPS_REG
    else:
        nn = n
        J = rec[L.off_jc:L.off_jc + 12 * nn].reshape(2, 6, nn)
        M = unpack_lower(rec[L.off_M:L.off_M + nn * (nn + 1) // 2], nn)
        h = rec[L.off_h:L.off_h + nn]
        F = rec[L.off_fee:L.off_fee + 12].reshape(2, 6)
        Minv = np.linalg.inv(M)
        A0 = np.vstack([(J[t] @ Minv)[:3] for t in range(2)])          # CartesianImpedanceCtrl rows 0..2
        if level == 0:
            A = A0; b = np.concatenate([(J[t] @ Minv)[:3] @ (J[t].T @ F[t]) for t in range(2)])
            eps = desc.eps_regularisation * QPOASES_EPS_REG
        elif desc.flags & FLAG_ELBOW_TASKS:
            # (_ee_task_right + _ee_task_left) / (_elbow_task_left + _elbow_task_right)  (QPPVMPlugin.cpp:154-166, 177-178)
            Je = rec[L.off_jelbow:L.off_jelbow + 12 * nn].reshape(2, 6, nn)
            Fe = rec[L.off_felbow:L.off_felbow + 12].reshape(2, 6)
            A = np.vstack([(Je[t] @ Minv)[:3] for t in range(2)])
            b = np.concatenate([(Je[t] @ Minv)[:3] @ (Je[t].T @ Fe[t]) for t in range(2)])
            eps = desc.eps_regularisation * QPOASES_EPS_REG
        else:
            A = Minv; b = Minv @ rec[L.off_tauj:L.off_tauj + nn]        # JointImpedanceCtrl
            eps = 0.0
        tl = rec[L.off_taulim:L.off_taulim + 2 * nn]
        blo, bhi = tl[:nn] - h, tl[nn:] - h                            # TorqueLimits (QPPVMPlugin.cpp:203-205)
        if desc.flags & FLAG_JOINT_LIMITS:                             # torque::JointLimits: bounds on the same variable
            jl = rec[L.off_jlim:L.off_jlim + 2 * nn]
            blo, bhi = np.maximum(blo, jl[:nn]), np.minimum(bhi, jl[nn:])
        rows, lo, hi = [np.eye(nn)], [blo], [bhi]
        if level == 1:
            rows.append(A0); lo.append(A0 @ x0); hi.append(A0 @ x0)
    return A, b, np.vstack(rows), np.concatenate(lo), np.concatenate(hi), eps
