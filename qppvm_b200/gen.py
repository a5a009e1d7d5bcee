"""Deterministic synthetic humanoid states -> QP records (SURVEY.md 8(d) "Synthetic inputs").

The reference gets J, Jdot*qdot, M, h from XBot::ModelInterface after
``_model->update()`` (ref:src/ForceAcc.cpp:256-282, ref:src/QPPVMPlugin.cpp:344-353) and
builds the task right-hand sides inside OpenSoT.  That producer is not
part of the hot path; this module stands in for it: a fixed kinematic tree (floating
pelvis, legs 2x6, waist 3, arms 2x7 [+ neck 2 + hands 2]) evaluated in FP64 with batched
numpy, written straight into the record layout of ``include/qppvm_b200.h``.

Determinism / sharding: problem ``i`` draws its DRAWS uniforms from Philox(key=seed) advanced
to ``i*DRAWS/4`` so any rank can regenerate exactly its slice ``[start, start+count)``.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtri

from .layout import (Desc, Layout, layout, KIND_FORCEACC, KIND_TORQUE, FLAG_FRICTION_CONES, FLAG_FULL_WRENCH, FLAG_ELBOW_TASKS, FLAG_JOINT_LIMITS, FLAG_COM_TASK,
                     FLAG_TORQUE_LIMITS)

DRAWS = 256            # uniforms reserved per problem (multiple of 4: Philox yields 4 per step)
BASE_SEED = 20260118   # seed = 20260118 + 1000 * config_index
GRAVITY = 9.81

_AX = {"x": (1.0, 0.0, 0.0), "y": (0.0, 1.0, 0.0), "z": (0.0, 0.0, 1.0)}


class Robot:
    """Fixed kinematic tree.  Body 0 = pelvis (floating base); body k>=1 moves with joint k-1."""

    def __init__(self, n_a: int):
        # (name, parent body, axis, offset from parent frame [m], mass [kg], com [m], tau_max [Nm])
        J = []

        def chain(prefix, parent, specs):
            p = parent
            for (nm, ax, off, m, com, tmax) in specs:
                J.append((prefix + nm, p, ax, off, m, com, tmax))
                p = len(J)          # body index of the link just added
            return p

        heavy = n_a >= 33           # WALK-MAN-like (~120 kg) vs COMAN-like (~30 kg)
        s = 4.0 if heavy else 1.0   # mass scale
        ln = 1.25 if heavy else 1.0  # length scale
        legs = lambda sy: [
            ("hip_pitch", "y", (0.0, sy * 0.07 * ln, -0.05 * ln), 1.2 * s, (0, 0, -0.02), 120 * s),
            ("hip_roll", "x", (0.0, 0.0, 0.0), 0.8 * s, (0, 0, -0.03), 100 * s),
            ("hip_yaw", "z", (0.0, 0.0, -0.05 * ln), 2.2 * s, (0, 0, -0.10 * ln), 60 * s),
            ("knee", "y", (0.0, 0.0, -0.22 * ln), 1.6 * s, (0, 0, -0.10 * ln), 120 * s),
            ("ank_pitch", "y", (0.0, 0.0, -0.22 * ln), 0.5 * s, (0, 0, -0.01), 80 * s),
            ("ank_roll", "x", (0.0, 0.0, 0.0), 0.7 * s, (0.02, 0, -0.04 * ln), 60 * s),
        ]
        arms = lambda sy: [
            ("sh_pitch", "y", (0.0, sy * 0.15 * ln, 0.20 * ln), 0.9 * s, (0, sy * 0.02, 0), 60 * s),
            ("sh_roll", "x", (0.0, 0.0, 0.0), 0.6 * s, (0, 0, -0.03), 60 * s),
            ("sh_yaw", "z", (0.0, 0.0, -0.04 * ln), 1.0 * s, (0, 0, -0.08 * ln), 40 * s),
            ("elbow", "y", (0.0, 0.0, -0.18 * ln), 0.7 * s, (0, 0, -0.06 * ln), 40 * s),
            ("fore_yaw", "z", (0.0, 0.0, -0.05 * ln), 0.5 * s, (0, 0, -0.05 * ln), 20 * s),
            ("wr_pitch", "y", (0.0, 0.0, -0.12 * ln), 0.3 * s, (0, 0, -0.01), 15 * s),
            ("wr_roll", "x", (0.0, 0.0, 0.0), 0.4 * s, (0, 0, -0.04 * ln), 15 * s),
        ]
        self.foot = [chain("l_", 0, legs(+1.0)), chain("r_", 0, legs(-1.0))]
        torso = chain("waist_", 0, [
            ("roll", "x", (0.0, 0.0, 0.08 * ln), 0.8 * s, (0, 0, 0.02), 80 * s),
            ("pitch", "y", (0.0, 0.0, 0.0), 0.8 * s, (0, 0, 0.03), 120 * s),
            ("yaw", "z", (0.0, 0.0, 0.05 * ln), 7.0 * s, (0, 0, 0.12 * ln), 80 * s),
        ])
        self.hand = [chain("l_", torso, arms(+1.0)), chain("r_", torso, arms(-1.0))]
        if n_a == 33:
            chain("neck_", torso, [
                ("yaw", "z", (0.0, 0.0, 0.28 * ln), 0.3 * s, (0, 0, 0.02), 10 * s),
                ("pitch", "y", (0.0, 0.0, 0.03), 1.0 * s, (0, 0, 0.06), 10 * s),
            ])
            for k, sy in enumerate((+1.0, -1.0)):
                chain("lr"[k] + "_hand_", self.hand[k], [
                    ("grasp", "y", (0.0, 0.0, -0.06 * ln), 0.3 * s, (0, 0, -0.03), 8 * s)])
        elif n_a != 29:
            # generic limb padding for other shapes: extra single-joint links on the torso
            while len(J) < n_a:
                chain("aux%d_" % len(J), torso, [
                    ("j", "xyz"[len(J) % 3], (0.05, 0.0, 0.1), 0.4 * s, (0, 0, 0.03), 20 * s)])
        if len(J) != n_a:
            J = J[:n_a]
            self.foot = [min(f, n_a) for f in self.foot]
            self.hand = [min(h, n_a) for h in self.hand]
        self.n_a = n_a
        self.n_v = n_a + 6
        self.n_b = n_a + 1
        self.names = ["pelvis"] + [j[0] for j in J]
        self.parent = np.array([-1] + [j[1] for j in J])
        self.axis = np.array([(0.0, 0.0, 0.0)] + [_AX[j[2]] for j in J])
        self.offset = np.array([(0.0, 0.0, 0.0)] + [j[3] for j in J], dtype=np.float64)
        self.mass = np.array([(6.0 * s)] + [j[4] for j in J])
        self.com = np.array([(0.0, 0.0, 0.03)] + [j[5] for j in J], dtype=np.float64)
        # box-like inertia about the COM: I = m * r^2 * (1, 1, 0.5), r ~ 6 cm * length scale
        r2 = (0.06 * ln) ** 2
        self.inertia = np.stack([self.mass * r2, self.mass * r2, 0.5 * self.mass * r2], axis=1)
        self.tau_max = np.array([j[6] for j in J], dtype=np.float64)
        # "home": slightly bent knees / elbows
        qh = np.zeros(n_a)
        for i, j in enumerate(J):
            nm = j[0]
            if nm.endswith("hip_pitch"): qh[i] = -0.35
            elif nm.endswith("knee"): qh[i] = 0.70
            elif nm.endswith("ank_pitch"): qh[i] = -0.35
            elif nm.endswith("sh_pitch"): qh[i] = 0.35
            elif nm.endswith("sh_roll"): qh[i] = 0.15 if nm.startswith("l_") else -0.15
            elif nm.endswith("elbow"): qh[i] = -0.9
        self.q_home = qh
        # ancestors mask: anc[i, k] = 1 if joint k (body k+1) is on the path pelvis -> body i
        anc = np.zeros((self.n_b, n_a), dtype=bool)
        for i in range(1, self.n_b):
            b = i
            while b > 0:
                anc[i, b - 1] = True
                b = self.parent[b]
        self.anc = anc

    # ---- batched kinematics / dynamics ------------------------------------------------
    @staticmethod
    def _rot(axis, ang):
        """Rodrigues, axis (3,) unit, ang (B,) -> (B,3,3)."""
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        s, c = np.sin(ang)[:, None, None], np.cos(ang)[:, None, None]
        return np.eye(3)[None] + s * K[None] + (1 - c) * (K @ K)[None]

    def dynamics(self, q, qd, R0, p0, v0, w0, links):
        """q, qd (B,n_a); base rotation R0 (B,3,3), position p0, linear/angular velocity v0, w0 (B,3)
        (world frame).  Generalised velocity = [v0, w0, qd] (world-aligned 'mixed' convention).
        Returns dict with M (B,nv,nv), h (B,nv) and per requested link: J (B,6,nv), Jdqd (B,6), R (B,3,3).
        """
        B, nb, nv, na = q.shape[0], self.n_b, self.n_v, self.n_a
        R = np.empty((B, nb, 3, 3)); p = np.empty((B, nb, 3)); z = np.zeros((B, nb, 3))
        w = np.empty((B, nb, 3)); al = np.zeros((B, nb, 3)); a = np.zeros((B, nb, 3))
        R[:, 0], p[:, 0], w[:, 0] = R0, p0, w0
        for i in range(1, nb):
            pa = self.parent[i]
            r = np.einsum("bij,j->bi", R[:, pa], self.offset[i])
            p[:, i] = p[:, pa] + r
            z[:, i] = np.einsum("bij,j->bi", R[:, pa], self.axis[i])
            R[:, i] = R[:, pa] @ self._rot(self.axis[i], q[:, i - 1])
            zq = z[:, i] * qd[:, i - 1:i]
            w[:, i] = w[:, pa] + zq
            al[:, i] = al[:, pa] + np.cross(w[:, pa], zq)
            a[:, i] = a[:, pa] + np.cross(al[:, pa], r) + np.cross(w[:, pa], np.cross(w[:, pa], r))

        def jac(body, point):
            """Jacobian of world point `point` (B,3) rigidly attached to `body`: (B,6,nv)."""
            Jm = np.zeros((B, 6, nv))
            Jm[:, 0, 0] = Jm[:, 1, 1] = Jm[:, 2, 2] = 1.0
            Jm[:, 3, 3] = Jm[:, 4, 4] = Jm[:, 5, 5] = 1.0
            d = point - p[:, 0]
            Jm[:, 0, 4], Jm[:, 0, 5] = d[:, 2], -d[:, 1]          # -[d]x
            Jm[:, 1, 3], Jm[:, 1, 5] = -d[:, 2], d[:, 0]
            Jm[:, 2, 3], Jm[:, 2, 4] = d[:, 1], -d[:, 0]
            ks = np.nonzero(self.anc[body])[0]
            zz = z[:, ks + 1]                                      # (B,k,3)
            rr = point[:, None, :] - p[:, ks + 1]
            Jm[:, 0:3, 6 + ks] = np.cross(zz, rr).transpose(0, 2, 1)
            Jm[:, 3:6, 6 + ks] = zz.transpose(0, 2, 1)
            return Jm

        M = np.zeros((B, nv, nv)); h = np.zeros((B, nv))
        gvec = np.array([0.0, 0.0, GRAVITY])
        for i in range(nb):
            c = np.einsum("bij,j->bi", R[:, i], self.com[i])
            Jc = jac(i, p[:, i] + c)
            Jv, Jw = Jc[:, 0:3], Jc[:, 3:6]
            Iw = (R[:, i] * self.inertia[i][None, None, :]) @ R[:, i].transpose(0, 2, 1)
            JvT, JwT = Jv.transpose(0, 2, 1), Jw.transpose(0, 2, 1)
            M += self.mass[i] * (JvT @ Jv)                     # batched BLAS (einsum was 75 % of the generator)
            M += JwT @ (Iw @ Jw)
            ac = a[:, i] + np.cross(al[:, i], c) + np.cross(w[:, i], np.cross(w[:, i], c))
            # the base itself translates: its classical acceleration bias is zero (v0 is world-frame)
            h += (JvT @ (self.mass[i] * (ac + gvec))[:, :, None])[:, :, 0]
            Iww = (Iw @ w[:, i][:, :, None])[:, :, 0]
            h += (JwT @ ((Iw @ al[:, i][:, :, None])[:, :, 0] + np.cross(w[:, i], Iww))[:, :, None])[:, :, 0]
        out = {"M": 0.5 * (M + M.transpose(0, 2, 1)), "h": h, "links": {}}
        for b in links:
            out["links"][b] = dict(J=jac(b, p[:, b]), Jdqd=np.concatenate([a[:, b], al[:, b]], axis=1),
                                   R=R[:, b].copy(), p=p[:, b].copy())
        return out


_ROBOTS: dict[int, Robot] = {}


def robot_for(n_a: int) -> Robot:
    if n_a not in _ROBOTS:
        _ROBOTS[n_a] = Robot(n_a)
    return _ROBOTS[n_a]


def _uniforms(seed: int, start: int, count: int) -> np.ndarray:
    bg = np.random.Philox(key=seed)
    bg.advance(start * (DRAWS // 4))
    return np.random.Generator(bg).random((count, DRAWS))


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.empty(r.shape + (3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = -sp, cp * sr, cp * cr
    return R


def pack_lower(M: np.ndarray) -> np.ndarray:
    n = M.shape[-1]
    i, j = np.tril_indices(n)
    return M[..., i, j]


def unpack_lower(Mp: np.ndarray, n: int) -> np.ndarray:
    M = np.zeros(Mp.shape[:-1] + (n, n))
    i, j = np.tril_indices(n)
    M[..., i, j] = Mp
    M[..., j, i] = Mp
    return M


def state_doubles(desc: Desc) -> int:
    """Compact synthetic state of one problem (what the rigid-body front end consumes):
    q | qd | R0 (3x3 row-major) | p0 | base twist (v0, w0) | gains 4 | ori_err 3 | foot_err 6c | mu c | tau_scale n_a |
    waist position error 3."""
    na, c = desc.n_a, desc.n_contacts
    return 2 * na + 9 + 3 + 6 + 4 + 3 + 6 * c + c + na + 3


def state_offsets(desc: Desc) -> dict:
    na, c = desc.n_a, desc.n_contacts
    o, out = 0, {}
    for name, n in (("q", na), ("qd", na), ("R0", 9), ("p0", 3), ("tw", 6), ("gains", 4), ("ori_err", 3),
                    ("foot_err", 6 * c), ("mu", c), ("tau_scale", na), ("waist_pos_err", 3)):
        out[name] = (o, o + n)
        o += n
    return out


def generate_states(desc: Desc, count: int, seed: int = BASE_SEED, start: int = 0) -> np.ndarray:
    """Compact states for problems [start, start+count): float64 (count, state_doubles)."""
    na, c = desc.n_a, desc.n_contacts
    U = _uniforms(seed, start, count)
    k = 0

    def take(m):
        nonlocal k
        v = U[:, k:k + m]
        k += m
        return v

    def normal(m, sigma):
        return sigma * ndtri(np.clip(take(m), 1e-12, 1 - 1e-12))

    rob = robot_for(na)
    q = rob.q_home[None] + (take(na) * 0.6 - 0.3)             # q_home + U(-0.3, 0.3)
    qd = normal(na, 0.5)                                      # N(0, 0.5^2) rad/s
    rpy = take(3) * np.array([0.4, 0.4, 2 * np.pi]) - np.array([0.2, 0.2, np.pi])
    R0 = _rpy(rpy[:, 0], rpy[:, 1], rpy[:, 2])
    p0 = np.concatenate([take(2) * 2 - 1, 0.55 + 0.1 * take(1)], axis=1)
    tw = normal(6, 0.2)
    gains = 0.8 + 0.4 * take(4)                               # "perturbed gains" x U(0.8, 1.2)
    ori_err = take(3) * 0.1 - 0.05
    foot_err = normal(24, 1e-3)[:, :6 * c]
    mu = (0.4 + 0.5 * take(4))[:, :c]
    tau_scale = 0.5 + 0.5 * take(na)
    assert k <= DRAWS
    if desc.kind == KIND_TORQUE:                              # fixed base
        R0 = np.tile(np.eye(3), (count, 1, 1)); p0 = np.zeros((count, 3)); tw = np.zeros((count, 6))
    # waist position error = reference - current: the reference is captured once as "initial - 0.1 z"
    # (ref:src/ForceAcc.cpp:158-164,181), so at the first tick the error is (0, 0, -0.1); it then shrinks as the robot
    # moves (integrate_states with the tick's records)
    wpos = np.tile(np.array([0.0, 0.0, -0.1]), (count, 1))
    return np.concatenate([q, qd, R0.reshape(count, 9), p0, tw, gains, ori_err, foot_err, mu, tau_scale, wpos], axis=1)


def generate(desc: Desc, count: int, seed: int = BASE_SEED, start: int = 0,
             chunk: int = 2048) -> np.ndarray:
    """Records for problems [start, start+count): float64 array (count, rec_doubles)."""
    L = layout(desc)
    out = np.zeros((count, L.rec_doubles))
    for c0 in range(0, count, chunk):
        n = min(chunk, count - c0)
        out[c0:c0 + n] = records_from_states(desc, generate_states(desc, n, seed, start + c0))
    return out


def records_from_states(desc: Desc, states: np.ndarray) -> np.ndarray:
    """Rigid-body dynamics + task right-hand sides: compact states -> QP records (the CPU statement of what
    ``model->update()`` + the OpenSoT task updates produce each tick, and of the on-device front end)."""
    L = layout(desc)
    rob = robot_for(desc.n_a)
    na, nv, c = L.n_a, desc.n_a + 6, L.n_c
    n = states.shape[0]
    so = state_offsets(desc)
    g = lambda name: states[:, so[name][0]:so[name][1]]
    q, qd, p0, tw, gains, ori_err, foot_err, mu, tau_scale, wpos = (g(k) for k in
        ("q", "qd", "p0", "tw", "gains", "ori_err", "foot_err", "mu", "tau_scale", "waist_pos_err"))
    R0 = g("R0").reshape(n, 3, 3)

    rec = np.zeros((n, L.rec_doubles))
    if desc.kind == KIND_FORCEACC:
        contact_bodies = (rob.foot + rob.hand)[:c]
        dyn = rob.dynamics(q, qd, R0, p0, tw[:, 0:3], tw[:, 3:6], [0] + contact_bodies)
        v = np.concatenate([tw, qd], axis=1)
        lam_w, lam2_w = 100.0 * gains[:, 0:1], 20.0 * gains[:, 1:2]
        lam_p, lam2_p = 100.0 * gains[:, 2:3], 20.0 * gains[:, 3:4]
        Jw = dyn["links"][0]["J"]
        rec[:, L.off_jwaist:L.off_jwaist + 6 * nv] = Jw.reshape(n, -1)
        rec[:, L.off_M:L.off_M + nv * (nv + 1) // 2] = pack_lower(dyn["M"])
        rec[:, L.off_h:L.off_h + nv] = dyn["h"]
        rec[:, L.off_jdqd:L.off_jdqd + 6] = dyn["links"][0]["Jdqd"]
        # waist: position reference = current - 0.1 z (ref:src/ForceAcc.cpp:181), velocity ref 0
        e_w = np.concatenate([wpos, ori_err], axis=1)
        rec[:, L.off_rhs:L.off_rhs + 6] = lam_w * e_w - lam2_w * np.einsum("bij,bj->bi", Jw, v)
        for ci, b in enumerate(contact_bodies):
            Jc = dyn["links"][b]["J"]
            o = L.off_jc + ci * 6 * nv
            rec[:, o:o + 6 * nv] = Jc.reshape(n, -1)
            rec[:, L.off_jdqd + 6 * (1 + ci):L.off_jdqd + 6 * (2 + ci)] = dyn["links"][b]["Jdqd"]
            e_c = foot_err[:, 6 * ci:6 * ci + 6]
            rec[:, L.off_rhs + 6 * (1 + ci):L.off_rhs + 6 * (2 + ci)] = (
                lam_p * e_c - lam2_p * np.einsum("bij,bj->bi", Jc, v))
            if desc.flags & FLAG_FRICTION_CONES:
                o = L.off_cone + 10 * ci
                rec[:, o:o + 9] = dyn["links"][b]["R"].reshape(n, 9)
                rec[:, o + 9] = mu[:, ci]
            if desc.flags & FLAG_FULL_WRENCH:                    # "put 6 for full wrench" (ForceAcc.cpp:67): lb / ub of :75-76 in full
                o = L.off_fbox + 12 * ci
                rec[:, o:o + 12] = np.array([-1000.0, -1000.0, 10.0, -1.0, -1.0, -1.0, 1000.0, 1000.0, 1000.0, 1.0, 1.0, 1.0])
            else:
                o = L.off_fbox + 6 * ci
                rec[:, o:o + 6] = np.array([-1000.0, -1000.0, 10.0, 1000.0, 1000.0, 1000.0])  # ForceAcc.cpp:75-76
        # postural: qddot = l2 (0 - qdot) + l (q_home - q); base rows carry the damping term only
        e_p = np.concatenate([np.zeros((n, 6)), rob.q_home[None] - q], axis=1)
        o = L.off_rhs + 6 * (1 + c)
        rec[:, o:o + nv] = lam_p * e_p - lam2_p * v
        if desc.flags & FLAG_TORQUE_LIMITS:
            tmax = rob.tau_max[None] * tau_scale
            rec[:, L.off_taulim:L.off_taulim + na] = -tmax
            rec[:, L.off_taulim + na:L.off_taulim + 2 * na] = tmax
        if desc.flags & FLAG_COM_TASK:
            # OpenSoT tasks::force::CoM (ref:src/ForceAcc.cpp:103): centroidal dynamics on the contact wrenches,
            #   sum f_i = m (a_ref + l2 (0 - cdot) + l (c_ref - c)) + m g z,   sum (p_i - c) x f_i (+ tau_i) = -k_L L_c.
            # Total mass, centre of mass and momentum come out of the joint-space quantities the record already holds:
            # M[0:3, 0:3] = m I, M[0:3, 3:6] = -m [c - p0]x (mixed velocity convention), momentum = M[0:6] v.
            wd = 6 if desc.flags & FLAG_FULL_WRENCH else 3
            Mm = dyn["M"]
            m = Mm[:, 0, 0]
            dcom = np.stack([Mm[:, 1, 5], Mm[:, 2, 3], Mm[:, 0, 4]], axis=1) / m[:, None]      # c - p0
            mom = np.einsum("bij,bj->bi", Mm[:, :6], v)
            cdot = mom[:, :3] / m[:, None]
            Lc = mom[:, 3:6] - np.cross(dcom, mom[:, :3])          # angular momentum about the centre of mass
            A = np.zeros((n, 6, wd * c))
            for ci, b in enumerate(contact_bodies):
                r = dyn["links"][b]["p"] - (p0 + dcom)
                A[:, 0, wd * ci + 0] = A[:, 1, wd * ci + 1] = A[:, 2, wd * ci + 2] = 1.0
                A[:, 3, wd * ci + 1], A[:, 3, wd * ci + 2] = -r[:, 2], r[:, 1]             # [r]x
                A[:, 4, wd * ci + 0], A[:, 4, wd * ci + 2] = r[:, 2], -r[:, 0]
                A[:, 5, wd * ci + 0], A[:, 5, wd * ci + 1] = -r[:, 1], r[:, 0]
                if wd == 6:
                    A[:, 3, wd * ci + 3] = A[:, 4, wd * ci + 4] = A[:, 5, wd * ci + 5] = 1.0
            e_com = np.concatenate([ori_err[:, :2] * 0.4, -0.02 * np.ones((n, 1))], axis=1)    # c_ref - c
            b_lin = m[:, None] * (lam_p * e_com - lam2_p * cdot) + m[:, None] * np.array([0.0, 0.0, GRAVITY])
            b_ang = -10.0 * gains[:, 3:4] * Lc
            rec[:, L.off_com:L.off_com + 6 * wd * c] = A.reshape(n, -1)
            rec[:, L.off_com + 6 * wd * c:L.off_com + 6 * wd * c + 6] = np.concatenate([b_lin, b_ang], axis=1)
    elif desc.kind == KIND_TORQUE:
        # fixed base: drop the 6 base columns/rows of the floating-base quantities
        Z3 = np.zeros((n, 3))
        elbows = [hb - 3 for hb in rob.hand]                   # "arm1_4" / "arm2_4": fourth body of each seven-joint arm chain
        dyn = rob.dynamics(q, qd, np.tile(np.eye(3), (n, 1, 1)), Z3, Z3, Z3, rob.hand[::-1] + elbows)
        Mj = dyn["M"][:, 6:, 6:]
        rec[:, L.off_M:L.off_M + na * (na + 1) // 2] = pack_lower(Mj)
        rec[:, L.off_h:L.off_h + na] = dyn["h"][:, 6:]
        ferr = np.zeros((n, 12)); ferr[:, :foot_err.shape[1]] = foot_err[:, :12]
        for ti, b in enumerate(rob.hand[::-1]):                # right first (QPPVMPlugin.cpp:177)
            Jh = dyn["links"][b]["J"][:, :, 6:]
            rec[:, L.off_jc + ti * 6 * na:L.off_jc + (ti + 1) * 6 * na] = Jh.reshape(n, -1)
            e = np.concatenate([ferr[:, 6 * ti:6 * ti + 3] * 30.0, ori_err * (1 - 2 * ti)], axis=1)
            F = 700.0 * gains[:, 0:1] * e - 70.0 * gains[:, 1:2] * np.einsum("bij,bj->bi", Jh, qd)
            rec[:, L.off_fee + 6 * ti:L.off_fee + 6 * ti + 6] = F     # K=700, D=70: QPPVMPlugin.cpp:136-137
        rec[:, L.off_tauj:L.off_tauj + na] = (5.0 * gains[:, 2:3] * (rob.q_home[None] - q)
                                              - 2.0 * gains[:, 3:4] * qd)  # K=5, D=2: QPPVMPlugin.cpp:105-106
        tmax = rob.tau_max[None] * tau_scale
        rec[:, L.off_taulim:L.off_taulim + na] = -tmax
        rec[:, L.off_taulim + na:L.off_taulim + 2 * na] = tmax
        if desc.flags & FLAG_JOINT_LIMITS:
            # OpenSoT constraints::torque::JointLimits (ref:src/QPPVMPlugin.cpp:169-171): tau in [k (q_min - q) - d qdot,
            # k (q_max - q) - d qdot] (gains through setGains, :170; k0 / d0 are not defined in the reference: synthetic
            # k = 50, d = 4 here); joint range q_home +- 0.35 rad shrunk by 10 % on both sides (:119-122), so that some of
            # the states (q_home + U(-0.3, 0.3)) sit close enough to a limit for the bound to bind
            qmin, qmax = rob.q_home[None] - 0.35 * 0.8, rob.q_home[None] + 0.35 * 0.8
            kj, dj = 50.0 * gains[:, 2:3], 4.0 * gains[:, 3:4]
            rec[:, L.off_jlim:L.off_jlim + na] = kj * (qmin - q) - dj * qd
            rec[:, L.off_jlim + na:L.off_jlim + 2 * na] = kj * (qmax - q) - dj * qd
        if desc.flags & FLAG_ELBOW_TASKS:
            # elbow_left + elbow_right: CartesianImpedanceCtrl on "arm1_4" / "arm2_4" (ref:src/QPPVMPlugin.cpp:154-166),
            # default gains K = 100 I, D = I (these two tasks never get setStiffnessDamping)
            for ti, hb in enumerate(rob.hand):                 # left first (:178)
                Je = dyn["links"][hb - 3]["J"][:, :, 6:]
                rec[:, L.off_jelbow + ti * 6 * na:L.off_jelbow + (ti + 1) * 6 * na] = Je.reshape(n, -1)
                e = np.concatenate([ferr[:, 6 * ti + 3:6 * ti + 6] * 20.0, -ori_err * (1 - 2 * ti)], axis=1)
                Fe = 100.0 * gains[:, 0:1] * e - 1.0 * gains[:, 1:2] * np.einsum("bij,bj->bi", Je, qd)
                rec[:, L.off_felbow + 6 * ti:L.off_felbow + 6 * ti + 6] = Fe
    return rec


def config_seed(config_index: int) -> int:
    return BASE_SEED + 1000 * config_index


def integrate_states(desc: Desc, states: np.ndarray, out: np.ndarray, dt: float, recs: np.ndarray | None = None) -> np.ndarray:
    """Numpy mirror of integrate_states_kernel (the integration ref:src/ForceAcc.cpp:225-226
    carries): one control period with the solved acceleration; failed solves leave their state untouched.
    `out` is the solver's output block viewed as float64 (B, out_doubles).
    With the tick's records the stored task errors (waist position / orientation, contact poses) follow the motion:
    the references were captured once (ref:src/ForceAcc.cpp:158-164,181), so e <- e - dt v_link - dt^2/2 a_link with
    v_link = J v recovered from the record's right-hand side (rhs = lambda e - lambda2 J v) and a_link = J qdd + Jdot qdot."""
    from .layout import layout
    L, o = layout(desc), state_offsets(desc)
    nv, na = L.n_v, desc.n_a
    st = states.copy()
    trailer = (L.n_x + na)
    ok = out[:, trailer:trailer + 1].copy().view(np.int32)[:, 0] == 0
    x = out[:, :nv]
    h2 = 0.5 * dt * dt
    tw = states[:, o["tw"][0]:o["tw"][1]]
    p0 = states[:, o["p0"][0]:o["p0"][1]] + dt * tw[:, :3] + h2 * x[:, :3]
    th = dt * tw[:, 3:] + h2 * x[:, 3:6]
    a2 = (th * th).sum(axis=1)
    ang = np.sqrt(a2)
    big = a2 > 1e-12
    safe = np.where(big, ang, 1.0)
    A = np.where(big, np.sin(safe) / safe, 1.0 - a2 / 6.0)
    Bc = np.where(big, (1.0 - np.cos(safe)) / np.where(big, a2, 1.0), 0.5 - a2 / 24.0)
    R = states[:, o["R0"][0]:o["R0"][1]].reshape(-1, 3, 3)
    thb = np.broadcast_to(th[:, None, :], R.shape)
    Rc = R.transpose(0, 2, 1)                                  # rows = columns of R
    k1 = np.cross(thb, Rc); k2 = np.cross(thb, k1)
    Rn = (Rc + A[:, None, None] * k1 + Bc[:, None, None] * k2).transpose(0, 2, 1)
    q = states[:, o["q"][0]:o["q"][1]]; qd = states[:, o["qd"][0]:o["qd"][1]]
    new = states.copy()
    new[:, o["p0"][0]:o["p0"][1]] = p0
    new[:, o["tw"][0]:o["tw"][1]] = tw + dt * x[:, :6]
    new[:, o["R0"][0]:o["R0"][1]] = Rn.reshape(-1, 9)
    new[:, o["q"][0]:o["q"][1]] = q + dt * qd + h2 * x[:, 6:nv]
    new[:, o["qd"][0]:o["qd"][1]] = qd + dt * x[:, 6:nv]
    if recs is not None:
        c = L.n_c
        gains = states[:, o["gains"][0]:o["gains"][1]]
        for t in range(1 + c):
            lam = 100.0 * gains[:, 0 if t == 0 else 2][:, None]
            lam2 = 20.0 * gains[:, 1 if t == 0 else 3][:, None]
            if t == 0:
                e = np.concatenate([states[:, o["waist_pos_err"][0]:o["waist_pos_err"][1]], states[:, o["ori_err"][0]:o["ori_err"][1]]], axis=1)
                J = recs[:, L.off_jwaist:L.off_jwaist + 6 * nv].reshape(-1, 6, nv)
            else:
                e = states[:, o["foot_err"][0] + 6 * (t - 1):o["foot_err"][0] + 6 * t]
                J = recs[:, L.off_jc + (t - 1) * 6 * nv:L.off_jc + t * 6 * nv].reshape(-1, 6, nv)
            rhs = recs[:, L.off_rhs + 6 * t:L.off_rhs + 6 * t + 6]
            vl = (lam * e - rhs) / lam2
            al = np.einsum("bij,bj->bi", J, x) + recs[:, L.off_jdqd + 6 * t:L.off_jdqd + 6 * t + 6]
            en = e - dt * vl - h2 * al
            if t == 0:
                new[:, o["waist_pos_err"][0]:o["waist_pos_err"][1]] = en[:, :3]
                new[:, o["ori_err"][0]:o["ori_err"][1]] = en[:, 3:]
            else:
                new[:, o["foot_err"][0] + 6 * (t - 1):o["foot_err"][0] + 6 * t] = en
    st[ok] = new[ok]
    return st
