/*
 * qppvm_oracle.c — CPU restatement of the reference's per-tick whole-body QP path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under qppvm_b200/ links, imports or executes
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in OpenSoT (ADVRHumanoids/OpenSoT,
 * no version pinned: ref:CMakeLists.txt:18) and qpOASES 3.x (transitive, unpinned;
 * ref:src/QPPVMPlugin.cpp:21), neither of which is under /root/reference nor installed
 * here, and the reference has no tests or golden vectors.  This file
 * restates their published semantics and is cross-checked in tests/
 * against HiGHS (scipy-bundled), brute-force active-set enumeration and its own KKT
 * certificate.  Every level QP is strictly convex (H + eps I), so a point that passes
 * the KKT certificate IS the solution qpOASES converges to (up to its 2.2e-7
 * termination tolerance).
 *
 * What follows the reference, by function:
 *   assemble_forceacc()  explicit A_l, b_l, C, lA, uA exactly as the OpenSoT stack built at
 *                        ref:src/ForceAcc.cpp:63-137 produces them each tick (:184):
 *                        variables :63-70, wrench = force / Zero(3) :81, wrench bounds :74-76
 *                        :91-95, postural :105-107, dyn-feas :109-114, waist :118-122,
 *                        stack "waist / (postural + feet) << dyn_feas << wb" :131-133.
 *   assemble_torque()    ref:src/QPPVMPlugin.cpp:112-179 (+ torque-limit shift :203-205).
 *   cascade()            QPOases_sot::solve (ref:src/ForceAcc.cpp:188-193,
 *                        ref:src/QPPVMPlugin.cpp:246): per level H = A^T A, g = -A^T b, global
 *                        constraints + optimality rows A_j x = A_j x_j*, eps-regularisation
 *                        (eps_regularisation * 2.221e-13) and numRegularisationSteps proximal
 *                        re-solves.
 *   recover()            ref:src/ForceAcc.cpp:196-219 (tau = M qdd + h - sum J^T w), resp.
 *                        ref:src/QPPVMPlugin.cpp:246-256 (tau_d = tau_qp + h; zero on failure).
 *   gi_solve()           dense dual active-set QP (Goldfarb & Idnani 1983) standing in for
 *                        qpOASES' online active-set solver: same problem class, same
 *                        unique minimiser.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/qppvm_b200.h"

#define ORACLE_FACTOR_CHOLESKY 0  /* form H = A^T A + eps I, Cholesky (reference numerics)      */
#define ORACLE_FACTOR_QR       1  /* Householder QR of [A ; sqrt(eps) I] (never forms A^T A)     */

/* ------------------------------------------------------------------ layout (own arithmetic) */
int oracle_layout(const qppvm_desc* d, qppvm_layout* L)
{
    memset(L, 0, sizeof(*L));
    L->off_jlim = L->off_jelbow = L->off_felbow = L->off_com = -1;
    if (d->n_a < 1 || d->n_a > 58) return 1;
    int off = 0, row = 0;
    if (d->kind == QPPVM_KIND_FORCEACC) {
        int c = d->n_contacts;
        if (c < 1 || c > 4) return 1;
        int cones = (d->flags & QPPVM_FLAG_FRICTION_CONES) != 0;
        int tl = (d->flags & QPPVM_FLAG_TORQUE_LIMITS) != 0;
        int wd = (d->flags & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3;   /* variables per contact (ref:src/ForceAcc.cpp:67) */
        L->n_a = d->n_a; L->n_v = d->n_a + 6; L->n_c = c; L->n_x = L->n_v + wd * c;
        if (L->n_x > 64) return 1;
        int nv = L->n_v;
        L->row_dyn = row; row += 6;
        L->row_box = row; row += 6 * c;
        L->row_cone = cones ? row : -1; row += cones ? 5 * c : 0;
        L->row_tau = tl ? row : -1; row += tl ? d->n_a : 0;
        L->row_opt = row; row += QPPVM_M0;
        L->off_jwaist = off; off += 6 * nv;
        L->off_jc = off; off += c * 6 * nv;
        L->off_M = off; off += nv * (nv + 1) / 2;
        L->off_h = off; off += nv;
        L->off_jdqd = off; off += 6 * (1 + c);
        L->off_rhs = off; off += 6 * (1 + c) + nv;
        L->off_taulim = tl ? off : -1; off += tl ? 2 * d->n_a : 0;
        L->off_cone = cones ? off : -1; off += cones ? 10 * c : 0;
        L->off_fbox = off; off += 2 * wd * c;
        if (d->flags & QPPVM_FLAG_COM_TASK) { L->off_com = off; off += 6 * wd * c + 6; }
        if (d->flags & ~(QPPVM_FLAG_FRICTION_CONES | QPPVM_FLAG_TORQUE_LIMITS | QPPVM_FLAG_FULL_WRENCH | QPPVM_FLAG_COM_TASK)) return 1;
        L->off_fee = L->off_tauj = -1;
    } else if (d->kind == QPPVM_KIND_TORQUE) {
        if (d->n_contacts != 2 || (d->flags & ~(QPPVM_FLAG_JOINT_LIMITS | QPPVM_FLAG_ELBOW_TASKS))) return 1;
        int n = d->n_a;
        L->n_a = L->n_v = L->n_x = n; L->n_c = 2;
        L->row_dyn = L->row_cone = L->row_tau = -1;
        L->row_box = 0; L->row_opt = n; row = n + QPPVM_M0;
        L->off_jwaist = -1;
        L->off_jc = off; off += 12 * n;
        L->off_M = off; off += n * (n + 1) / 2;
        L->off_h = off; off += n;
        L->off_jdqd = L->off_rhs = -1;
        L->off_fee = off; off += 12;
        L->off_tauj = off; off += n;
        L->off_taulim = off; off += 2 * n;
        L->off_cone = L->off_fbox = -1;
        if (d->flags & QPPVM_FLAG_JOINT_LIMITS) { L->off_jlim = off; off += 2 * n; }
        if (d->flags & QPPVM_FLAG_ELBOW_TASKS) { L->off_jelbow = off; off += 12 * n; L->off_felbow = off; off += 12; }
    } else return 1;
    if (row > 128) return 1;
    L->n_rows = row;
    L->rec_doubles = off + (off & 1);
    L->out_bytes = 8 * (L->n_x + L->n_a) + 32;
    L->diag_doubles = L->n_x + 2 * row + QPPVM_M0;
    return 0;
}

/* ------------------------------------------------------------------ dense level QP */
typedef struct {
    int n, m, nc;          /* variables, task rows, constraint rows (two-sided)        */
    double *A, *b;         /* m x n row-major, m                                        */
    double *C, *lA, *uA;   /* nc x n row-major, nc, nc                                  */
    double eps;            /* regularisation added to the diagonal of H (0: none)       */
} level_qp;

static double mget(const double* Mp, int i, int j) /* packed lower, row-major */
{
    return i >= j ? Mp[i * (i + 1) / 2 + j] : Mp[j * (j + 1) / 2 + i];
}

/* ForceAcc stack.  level 0: waist Cartesian; level 1: postural + contact Cartesian tasks. */
static void assemble_forceacc(const qppvm_desc* d, const qppvm_layout* L, const double* rec,
                              int level, const double* x0, level_qp* q)
{
    const int n = L->n_x, nv = L->n_v, c = L->n_c, na = L->n_a;
    const int wd = (d->flags & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3;   /* wrench variables per contact */
    q->n = n;
    q->eps = d->eps_regularisation * QPPVM_QPOASES_EPS_REG;
    /* H = A^T W A, g = -lambda A^T W b, W = w I per task: rows scaled by sqrt(w), rhs by lambda sqrt(w).
     * 0.0 = not set = 1.0 (qppvm_desc).  Postural: A = [I 0] on all n_v rows, or without the six base rows (A.6). */
    const double lam = d->lambda_solver != 0.0 ? d->lambda_solver : 1.0;
    const double sw0 = sqrt(d->task_weight[0] != 0.0 ? d->task_weight[0] : 1.0);
    const double sw1 = sqrt(d->task_weight[1] != 0.0 ? d->task_weight[1] : 1.0);
    const double sw2 = sqrt(d->task_weight[2] != 0.0 ? d->task_weight[2] : 1.0);
    const double* Jw = rec + L->off_jwaist;
    if (level == 0) {
        q->m = 6;
        memset(q->A, 0, sizeof(double) * q->m * n);
        for (int r = 0; r < 6; ++r) {
            for (int j = 0; j < nv; ++j) q->A[r * n + j] = sw0 * Jw[r * nv + j];
            q->b[r] = sw0 * lam * (rec[L->off_rhs + r] - rec[L->off_jdqd + r]);
        }
    } else {
        const int com = (d->flags & QPPVM_FLAG_COM_TASK) != 0;
        q->m = nv + 6 * c + (com ? 6 : 0);
        memset(q->A, 0, sizeof(double) * q->m * n);
        for (int i = 0; i < nv; ++i) {                      /* Postural: A = [I 0] */
            const double s = (d->postural_actuated_only && i < 6) ? 0.0 : sw1;
            q->A[i * n + i] = s;
            q->b[i] = s * lam * rec[L->off_rhs + 6 * (1 + c) + i];
        }
        for (int ci = 0; ci < c; ++ci)                      /* contact-link Cartesian tasks */
            for (int r = 0; r < 6; ++r) {
                int row = nv + 6 * ci + r;
                const double* J = rec + L->off_jc + (ci * 6 + r) * nv;
                for (int j = 0; j < nv; ++j) q->A[row * n + j] = sw2 * J[j];
                q->b[row] = sw2 * lam * (rec[L->off_rhs + 6 * (1 + ci) + r] - rec[L->off_jdqd + 6 * (1 + ci) + r]);
            }
        if (com)                                            /* tasks::force::CoM (ref:src/ForceAcc.cpp:103): rows on the wrench variables */
            for (int r = 0; r < 6; ++r) {
                int row = nv + 6 * c + r;
                for (int j = 0; j < wd * c; ++j) q->A[row * n + nv + j] = rec[L->off_com + r * wd * c + j];
                q->b[row] = lam * rec[L->off_com + 6 * wd * c + r];
            }
    }
    /* global constraints, in stack order */
    q->nc = (level == 0) ? L->row_opt : L->n_rows;
    memset(q->C, 0, sizeof(double) * q->nc * n);
    const double* Mp = rec + L->off_M;
    const double* h = rec + L->off_h;
    for (int r = 0; r < 6; ++r) {                           /* DynamicFeasibility: base rows */
        double* Cr = q->C + (L->row_dyn + r) * n;
        for (int j = 0; j < nv; ++j) Cr[j] = mget(Mp, r, j);
        for (int ci = 0; ci < c; ++ci)
            for (int k = 0; k < wd; ++k)                    /* wrench = [f;0]: linear rows only (all six with full wrenches) */
                Cr[nv + wd * ci + k] = -rec[L->off_jc + (ci * 6 + k) * nv + r];
        q->lA[L->row_dyn + r] = q->uA[L->row_dyn + r] = -h[r];
    }
    for (int ci = 0; ci < c; ++ci) {                        /* wrench bounds (GenericConstraint) */
        const double* fb = rec + L->off_fbox + 2 * wd * ci;
        for (int k = 0; k < 6; ++k) {
            int row = L->row_box + 6 * ci + k;
            if (k < wd) { q->C[row * n + nv + wd * ci + k] = 1.0; q->lA[row] = fb[k]; q->uA[row] = fb[wd + k]; }
            else { q->lA[row] = -1.0; q->uA[row] = 1.0; }   /* zero rows: torque part of the wrench = force / Zero(3) */
        }
    }
    if (L->row_cone >= 0)
        for (int ci = 0; ci < c; ++ci) {                    /* friction pyramid on R^T f (App. A.8) */
            const double* R = rec + L->off_cone + 10 * ci;
            const double mu = R[9] / sqrt(2.0);
            const double Ci[5][3] = {{1, 0, -mu}, {-1, 0, -mu}, {0, 1, -mu}, {0, -1, -mu}, {0, 0, -1}};
            for (int j = 0; j < 5; ++j) {
                int row = L->row_cone + 5 * ci + j;
                for (int k = 0; k < 3; ++k) {
                    double s = 0;
                    for (int m = 0; m < 3; ++m) s += Ci[j][m] * R[k * 3 + m];
                    q->C[row * n + nv + wd * ci + k] = s;
                }
                q->lA[row] = -QPPVM_INFTY; q->uA[row] = 0.0;
            }
        }
    if (L->row_tau >= 0)
        for (int a = 0; a < na; ++a) {                      /* torque limits (App. A.7) */
            int row = L->row_tau + a;
            double* Cr = q->C + row * n;
            for (int j = 0; j < nv; ++j) Cr[j] = mget(Mp, 6 + a, j);
            for (int ci = 0; ci < c; ++ci)
                for (int k = 0; k < wd; ++k)
                    Cr[nv + wd * ci + k] = -rec[L->off_jc + (ci * 6 + k) * nv + 6 + a];
            q->lA[row] = rec[L->off_taulim + a] - h[6 + a];
            q->uA[row] = rec[L->off_taulim + na + a] - h[6 + a];
        }
    if (level == 1)
        for (int r = 0; r < 6; ++r) {                       /* optimality rows of level 0 */
            int row = L->row_opt + r;
            double s = 0;
            for (int j = 0; j < nv; ++j) { q->C[row * n + j] = Jw[r * nv + j]; s += Jw[r * nv + j] * x0[j]; }
            q->lA[row] = q->uA[row] = s;
        }
}

/* in-place Cholesky of SPD matrix (row-major n x n), lower factor; returns 0 ok */
static int chol_lower(double* S, int n)
{
    for (int j = 0; j < n; ++j) {
        double s = S[j * n + j];
        for (int k = 0; k < j; ++k) s -= S[j * n + k] * S[j * n + k];
        if (!(s > 0.0)) return 1;
        double dj = sqrt(s);
        S[j * n + j] = dj;
        for (int i = j + 1; i < n; ++i) {
            double t = S[i * n + j];
            for (int k = 0; k < j; ++k) t -= S[i * n + k] * S[j * n + k];
            S[i * n + j] = t / dj;
        }
        for (int k = j + 1; k < n; ++k) S[j * n + k] = 0.0;
    }
    return 0;
}

/* Torque stack (fixed base).  level 0: two 3-row Cartesian impedance tasks; level 1: joint impedance. */
static int assemble_torque(const qppvm_desc* d, const qppvm_layout* L, const double* rec,
                           int level, const double* x0, level_qp* q, double* wk /* >= 3 n^2 */)
{
    const int n = L->n_x;
    q->n = n;
    double* Minv = wk;                 /* n x n */
    double* Lm = wk + n * n;           /* chol */
    double* T = wk + 2 * n * n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Lm[i * n + j] = mget(rec + L->off_M, i, j);
    if (chol_lower(Lm, n)) return 1;
    /* Minv = L^-T L^-1 : T = L^-1 (lower) */
    memset(T, 0, sizeof(double) * n * n);
    for (int c = 0; c < n; ++c) {
        for (int i = c; i < n; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) s -= Lm[i * n + k] * T[k * n + c];
            T[i * n + c] = s / Lm[i * n + i];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = (i > j ? i : j); k < n; ++k) s += T[k * n + i] * T[k * n + j];
            Minv[i * n + j] = s;
        }
    /* rows of J Minv for both hands (rows 0..2 are the stacked ones, OpenSoT::Indices::range(0,2)) */
    const int elbows = (d->flags & QPPVM_FLAG_ELBOW_TASKS) != 0;
    if (level == 0 || elbows) {
        q->m = 6;
        q->eps = d->eps_regularisation * QPPVM_QPOASES_EPS_REG;   /* CartesianImpedanceCtrl, HST_SEMIDEF: regularised */
    } else {
        q->m = n;
        q->eps = 0.0;                                             /* HST_POSDEF: not regularised */
    }
    double A0[6 * 64];
    for (int t = 0; t < 2; ++t) {
        const double* J = rec + L->off_jc + t * 6 * n;
        const double* F = rec + L->off_fee + 6 * t;
        double JtF[64];
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int r = 0; r < 6; ++r) s += J[r * n + j] * F[r];
            JtF[j] = s;
        }
        for (int r = 0; r < 3; ++r) {
            double* Ar = A0 + (3 * t + r) * n;
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int k = 0; k < n; ++k) s += J[r * n + k] * Minv[k * n + j];
                Ar[j] = s;
            }
            if (level == 0) {
                double s = 0;
                for (int j = 0; j < n; ++j) s += Ar[j] * JtF[j];    /* b = A J^T F (App. A.3) */
                q->b[3 * t + r] = s;
            }
        }
    }
    if (level == 0) memcpy(q->A, A0, sizeof(double) * 6 * n);
    else if (elbows) {
        /* level 1 = elbow_left + elbow_right (ref:src/QPPVMPlugin.cpp:154-166, stack of :177-178): the same task type as
         * the hands, A = (J M^-1)[0..2], b = A J^T F (App. A.3) */
        for (int t = 0; t < 2; ++t) {
            const double* J = rec + L->off_jelbow + t * 6 * n;
            const double* F = rec + L->off_felbow + 6 * t;
            double JtF[64];
            for (int j = 0; j < n; ++j) {
                double sj = 0;
                for (int r = 0; r < 6; ++r) sj += J[r * n + j] * F[r];
                JtF[j] = sj;
            }
            for (int r = 0; r < 3; ++r) {
                double* Ar = q->A + (3 * t + r) * n;
                double sb = 0;
                for (int j = 0; j < n; ++j) {
                    double sj = 0;
                    for (int k = 0; k < n; ++k) sj += J[r * n + k] * Minv[k * n + j];
                    Ar[j] = sj;
                }
                for (int j = 0; j < n; ++j) sb += Ar[j] * JtF[j];
                q->b[3 * t + r] = sb;
            }
        }
    } else {
        memcpy(q->A, Minv, sizeof(double) * n * n);                 /* A = M^-1, b = M^-1 tau_j */
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) s += Minv[i * n + j] * rec[L->off_tauj + j];
            q->b[i] = s;
        }
    }
    q->nc = (level == 0) ? n : n + QPPVM_M0;
    memset(q->C, 0, sizeof(double) * q->nc * n);
    const double* h = rec + L->off_h;
    for (int i = 0; i < n; ++i) {                                   /* TorqueLimits: tau_lim_const - h */
        q->C[i * n + i] = 1.0;
        q->lA[i] = rec[L->off_taulim + i] - h[i];
        q->uA[i] = rec[L->off_taulim + n + i] - h[i];
        if (d->flags & QPPVM_FLAG_JOINT_LIMITS) {
            /* torque::JointLimits (ref:src/QPPVMPlugin.cpp:169-171): simple bounds on the same variable; AutoStack::getBounds
             * intersects the simple bounds of all global constraints */
            q->lA[i] = fmax(q->lA[i], rec[L->off_jlim + i]);
            q->uA[i] = fmin(q->uA[i], rec[L->off_jlim + n + i]);
        }
    }
    if (level == 1)
        for (int r = 0; r < 6; ++r) {
            double s = 0;
            for (int j = 0; j < n; ++j) { q->C[(n + r) * n + j] = A0[r * n + j]; s += A0[r * n + j] * x0[j]; }
            q->lA[n + r] = q->uA[n + r] = s;
        }
    return 0;
}

/* ------------------------------------------------------------------ factor of H + eps I */
/* Produces J (n x n, row-major) with J^T (H + eps I) J = I, i.e. J = R^-1 for R^T R = H + eps I. */
static int factor_J(const level_qp* q, int mode, double* J, double* wk /* >= (m+n)*n + n*n */)
{
    const int n = q->n, m = q->m;
    double* R = wk;                     /* n x n upper */
    if (mode == ORACLE_FACTOR_CHOLESKY) {
        double* H = wk + n * n;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                double s = 0;
                for (int r = 0; r < m; ++r) s += q->A[r * n + i] * q->A[r * n + j];
                H[i * n + j] = s;
            }
        for (int i = 0; i < n; ++i) H[i * n + i] += q->eps;
        if (chol_lower(H, n)) return 1;                  /* H = L L^T, R = L^T */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) R[i * n + j] = (j >= i) ? H[j * n + i] : 0.0;
    } else {
        const int mm = m + n;
        double* S = wk + n * n;                          /* (m+n) x n stacked, row-major */
        memcpy(S, q->A, sizeof(double) * m * n);
        memset(S + m * n, 0, sizeof(double) * n * n);
        const double se = sqrt(q->eps);
        for (int i = 0; i < n; ++i) S[(m + i) * n + i] = se;
        for (int k = 0; k < n; ++k) {                    /* Householder, column k */
            double nrm = 0;
            for (int r = k; r < mm; ++r) nrm += S[r * n + k] * S[r * n + k];
            nrm = sqrt(nrm);
            if (nrm == 0.0) return 1;
            double alpha = S[k * n + k];
            double beta = (alpha >= 0) ? -nrm : nrm;
            double v0 = alpha - beta;
            double vtv = v0 * v0;
            for (int r = k + 1; r < mm; ++r) vtv += S[r * n + k] * S[r * n + k];
            for (int j = k + 1; j < n; ++j) {
                double s = v0 * S[k * n + j];
                for (int r = k + 1; r < mm; ++r) s += S[r * n + k] * S[r * n + j];
                s *= 2.0 / vtv;
                S[k * n + j] -= s * v0;
                for (int r = k + 1; r < mm; ++r) S[r * n + j] -= s * S[r * n + k];
            }
            S[k * n + k] = beta;
        }
        for (int i = 0; i < n; ++i) {
            double sg = (S[i * n + i] < 0) ? -1.0 : 1.0; /* make diag(R) > 0 */
            for (int j = 0; j < n; ++j) R[i * n + j] = (j >= i) ? sg * S[i * n + j] : 0.0;
        }
    }
    /* J = R^-1 (upper) by back substitution, column by column */
    memset(J, 0, sizeof(double) * n * n);
    for (int c = 0; c < n; ++c)
        for (int i = c; i >= 0; --i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k <= c; ++k) s -= R[i * n + k] * J[k * n + c];
            J[i * n + c] = s / R[i * n + i];
        }
    return 0;
}

/* ------------------------------------------------------------------ Goldfarb-Idnani */
typedef struct {
    int n, k;              /* k active constraints                                   */
    double* J;             /* n x n row-major, updated by the rotations               */
    double* R;             /* n x n row-major upper, leading k x k used               */
    double* u;             /* multipliers (>= 0 for inequalities)                     */
    int* row; int* sgn; int* iseq;
} gi_state;

static void givens(double a, double b, double* c, double* s)
{
    if (b == 0.0) { *c = 1.0; *s = 0.0; return; }
    double r = hypot(a, b);
    *c = a / r; *s = b / r;
}

static void gi_add(gi_state* g, double* d, int row, int sgn, int iseq, double u_new)
{
    const int n = g->n, k = g->k;
    for (int j = n - 1; j > k; --j) {       /* zero d[j] into d[j-1] */
        double c, s;
        givens(d[j - 1], d[j], &c, &s);
        if (s == 0.0) continue;
        d[j - 1] = c * d[j - 1] + s * d[j]; d[j] = 0.0;
        for (int i = 0; i < n; ++i) {
            double a = g->J[i * n + j - 1], b = g->J[i * n + j];
            g->J[i * n + j - 1] = c * a + s * b;
            g->J[i * n + j] = -s * a + c * b;
        }
    }
    for (int i = 0; i <= k; ++i) g->R[i * n + k] = d[i];
    g->u[k] = u_new; g->row[k] = row; g->sgn[k] = sgn; g->iseq[k] = iseq;
    g->k = k + 1;
}

static void gi_drop(gi_state* g, int l)
{
    const int n = g->n, k = g->k;
    for (int j = l; j < k - 1; ++j) {
        for (int i = 0; i < k; ++i) g->R[i * n + j] = g->R[i * n + j + 1];
        g->u[j] = g->u[j + 1]; g->row[j] = g->row[j + 1]; g->sgn[j] = g->sgn[j + 1]; g->iseq[j] = g->iseq[j + 1];
    }
    for (int i = l; i < k - 1; ++i) {       /* restore triangular form: kill R[i+1][i] */
        double c, s;
        givens(g->R[i * n + i], g->R[(i + 1) * n + i], &c, &s);
        if (s == 0.0) continue;
        for (int j = i; j < k - 1; ++j) {
            double a = g->R[i * n + j], b = g->R[(i + 1) * n + j];
            g->R[i * n + j] = c * a + s * b;
            g->R[(i + 1) * n + j] = -s * a + c * b;
        }
        for (int r = 0; r < n; ++r) {
            double a = g->J[r * n + i], b = g->J[r * n + i + 1];
            g->J[r * n + i] = c * a + s * b;
            g->J[r * n + i + 1] = -s * a + c * b;
        }
    }
    g->k = k - 1;
}

/* min 1/2 x^T G x + g^T x  s.t. lA <= C x <= uA, given J (J^T G J = I) and the unconstrained
 * minimiser x (in/out).  y[row] > 0: active at lA, < 0: active at uA.  Returns QPPVM_STATUS_*. */
static int gi_solve(int n, const double* Jfac, double* x, int nc, const double* C,
                    const double* lA, const double* uA, int max_iter, double* y, int* iters,
                    double* wk /* >= 2 n^2 + 4 n */, int* iwk /* >= 3 n + nc */,
                    const unsigned char* cand /* rows to try first (hot start), or NULL */)
{
    gi_state g;
    g.n = n; g.k = 0;
    g.J = wk; g.R = wk + n * n; g.u = wk + 2 * n * n;
    double* d = g.u + n; double* z = d + n; double* r = z + n;
    g.row = iwk; g.sgn = iwk + n; g.iseq = iwk + 2 * n;
    int* act = iwk + 3 * n;                /* 0 inactive, +1 lower, -1 upper, 2 dropped-redundant eq */
    memcpy(g.J, Jfac, sizeof(double) * n * n);
    memset(g.R, 0, sizeof(double) * n * n);
    memset(act, 0, sizeof(int) * nc);
    memset(y, 0, sizeof(double) * nc);
    int it = 0, status = QPPVM_STATUS_OK;
    int next_eq = 0;                       /* equalities are added first, in row order */
    for (;;) {
        /* ---- step 1: pick the constraint to add */
        int p = -1, psgn = 0, peq = 0; double sp = 0;
        while (next_eq < nc && !(lA[next_eq] == uA[next_eq])) ++next_eq;
        if (next_eq < nc) {
            p = next_eq; peq = 1;
            double cx = 0;
            for (int j = 0; j < n; ++j) cx += C[p * n + j] * x[j];
            double s = cx - lA[p];
            psgn = (s > 0) ? -1 : 1;
            sp = -fabs(s);
            ++next_eq;
        } else {
            double worst = 0, wsl = 0;
            for (int i = 0; i < nc; ++i) {
                /* the side opposite to an active one is still checked: an empty box (lA > uA) must surface as
                 * infeasible instead of being masked by the active side */
                if (act[i] == 2 || act[i] == 3 || lA[i] == uA[i]) continue;
                double cx = 0;
                for (int j = 0; j < n; ++j) cx += C[i * n + j] * x[j];
                double tol = 1e-9 * fmax(1.0, fabs(cx));
                /* hot start: violated rows of the previous working set go first (any violated row is a valid pivot) */
                const double pri = (cand && cand[i]) ? 0x1p100 : 1.0;
                if (act[i] != 1 && lA[i] > -0.5 * QPPVM_INFTY && cx - lA[i] < -tol && (cx - lA[i]) * pri < worst) { worst = (cx - lA[i]) * pri; wsl = cx - lA[i]; p = i; psgn = 1; }
                if (act[i] != -1 && uA[i] < 0.5 * QPPVM_INFTY && uA[i] - cx < -tol && (uA[i] - cx) * pri < worst) { worst = (uA[i] - cx) * pri; wsl = uA[i] - cx; p = i; psgn = -1; }
            }
            if (p < 0) break;              /* primal feasible: optimal */
            sp = wsl;
        }
        double up = 0.0;                   /* multiplier of p while it is being added */
        for (;;) {
            if (it >= max_iter) { status = QPPVM_STATUS_MAX_ITER; goto done; }
            /* ---- step 2: directions.  d = J^T n+, z = J2 d2, r = R^-1 d1 */
            const int k = g.k;
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int i = 0; i < n; ++i) s += g.J[i * n + j] * C[p * n + i];
                d[j] = psgn * s;
            }
            double dd = 0, d2 = 0;
            for (int j = 0; j < n; ++j) { dd += d[j] * d[j]; if (j >= k) d2 += d[j] * d[j]; }
            for (int i = 0; i < n; ++i) {
                double s = 0;
                for (int j = k; j < n; ++j) s += g.J[i * n + j] * d[j];
                z[i] = s;
            }
            for (int i = k - 1; i >= 0; --i) {
                double s = d[i];
                for (int j = i + 1; j < k; ++j) s -= g.R[i * n + j] * r[j];
                r[i] = s / g.R[i * n + i];
            }
            /* ---- step 3: step lengths */
            int dependent = !(d2 > 1e-22 * dd) || k >= n;
            double t1 = INFINITY; int l = -1;
            for (int i = 0; i < k; ++i)
                if (!g.iseq[i] && r[i] > 0 && g.u[i] / r[i] < t1) { t1 = g.u[i] / r[i]; l = i; }
            double t2 = dependent ? INFINITY : -sp / d2;
            if (peq && dependent && l < 0) {
                /* linearly dependent equality: redundant if consistent, else infeasible */
                if (-sp <= 1e-8 * fmax(1.0, fabs(lA[p]))) { act[p] = 2; break; }
                status = QPPVM_STATUS_INFEASIBLE; goto done;
            }
            if (t1 == INFINITY && t2 == INFINITY) {
                /* dependent on the working set and nothing can leave: an implied inequality whose violation is
                 * rounding noise (degenerate vertex: the optimality rows pin the previous level's optimum onto
                 * its active bounds) counts as satisfied within tolerance and is not scanned again */
                double bnd = psgn > 0 ? lA[p] : uA[p];
                if (!peq && -sp <= 1e-6 * fmax(1.0, fabs(bnd))) { act[p] = 3; break; }
                status = QPPVM_STATUS_INFEASIBLE; goto done;
            }
            double t = (t1 < t2) ? t1 : t2;
            /* ---- step 4: move */
            for (int i = 0; i < k; ++i) g.u[i] -= t * r[i];
            up += t;
            if (t2 != INFINITY) {
                for (int i = 0; i < n; ++i) x[i] += t * z[i];
                sp += t * d2;
            }
            ++it;
            if (t == t2) {                 /* full step: p becomes active */
                gi_add(&g, d, p, psgn, peq, up);
                act[p] = psgn;
                break;
            }
            act[g.row[l]] = 0;             /* partial step: drop blocking constraint l, retry p */
            gi_drop(&g, l);
        }
    }
done:
    for (int i = 0; i < g.k; ++i) y[g.row[i]] = g.sgn[i] * g.u[i];
    *iters = it;
    return status;
}

/* ------------------------------------------------------------------ KKT certificate */
static double kkt_residual(const level_qp* q, const double* g, const double* x, const double* y)
{
    const int n = q->n, m = q->m, nc = q->nc;
    double Ax[160];
    for (int r = 0; r < m; ++r) { double s = 0; for (int j = 0; j < n; ++j) s += q->A[r * n + j] * x[j]; Ax[r] = s; }
    double gmax = 0, hxmax = 0, xmax = 0, cxmax = 0, rs = 0, rp = 0, rc = 0, ymax = 0;
    for (int j = 0; j < n; ++j) {
        double hx = q->eps * x[j];
        for (int r = 0; r < m; ++r) hx += q->A[r * n + j] * Ax[r];
        double st = hx + g[j];
        for (int i = 0; i < nc; ++i) st -= q->C[i * n + j] * y[i];
        rs = fmax(rs, fabs(st)); gmax = fmax(gmax, fabs(g[j])); hxmax = fmax(hxmax, fabs(hx));
        xmax = fmax(xmax, fabs(x[j]));
    }
    for (int i = 0; i < nc; ++i) ymax = fmax(ymax, fabs(y[i]));
    for (int i = 0; i < nc; ++i) {
        double cx = 0;
        for (int j = 0; j < n; ++j) cx += q->C[i * n + j] * x[j];
        cxmax = fmax(cxmax, fabs(cx));
        double vl = q->lA[i] - cx, vu = cx - q->uA[i];
        rp = fmax(rp, fmax(0.0, fmax(vl, vu)));
        if (q->lA[i] == q->uA[i]) continue;
        if (y[i] > 0) rc = fmax(rc, y[i] * fabs(cx - q->lA[i]));
        if (y[i] < 0) rc = fmax(rc, -y[i] * fabs(q->uA[i] - cx));
    }
    rs /= fmax(1.0, fmax(gmax, hxmax));
    rp /= fmax(1.0, fmax(xmax, cxmax));
    rc /= fmax(1.0, ymax) * fmax(1.0, cxmax);
    return fmax(rs, fmax(rp, rc));
}

/* ------------------------------------------------------------------ one record */
typedef struct {
    double *A, *b, *C, *lA, *uA, *J, *wk, *g, *x, *xu, *y, *giwk;
    int* iwk;
} scratch;

static scratch* scratch_new(void)
{
    const int n = 64, m = 160, nc = 128;
    scratch* s = (scratch*)calloc(1, sizeof(scratch));
    s->A = (double*)malloc(sizeof(double) * m * n); s->b = (double*)malloc(sizeof(double) * m);
    s->C = (double*)malloc(sizeof(double) * nc * n);
    s->lA = (double*)malloc(sizeof(double) * nc); s->uA = (double*)malloc(sizeof(double) * nc);
    s->J = (double*)malloc(sizeof(double) * n * n);
    s->wk = (double*)malloc(sizeof(double) * ((m + n) * n + 4 * n * n));
    s->g = (double*)malloc(sizeof(double) * n); s->x = (double*)malloc(sizeof(double) * n);
    s->xu = (double*)malloc(sizeof(double) * n); s->y = (double*)malloc(sizeof(double) * nc);
    s->giwk = (double*)malloc(sizeof(double) * (2 * n * n + 4 * n));
    s->iwk = (int*)malloc(sizeof(int) * (3 * n + nc));
    return s;
}
static void scratch_free(scratch* s)
{
    free(s->A); free(s->b); free(s->C); free(s->lA); free(s->uA); free(s->J); free(s->wk);
    free(s->g); free(s->x); free(s->xu); free(s->y); free(s->giwk); free(s->iwk); free(s);
}

/* Solves one level: x (out), y (out, nc), returns status; *kkt = certificate of the solved problem. */
static int solve_level(const level_qp* q, int mode, int n_reg_steps, int max_iter, scratch* s,
                       double* x, double* y, int* iters, double* kkt, unsigned char* cand)
{
    const int n = q->n, m = q->m;
    if (factor_J(q, mode, s->J, s->wk)) return QPPVM_STATUS_NUMERIC;
    double g_orig[64];
    for (int j = 0; j < n; ++j) {                        /* g = -A^T b */
        double v = 0;
        for (int r = 0; r < m; ++r) v += q->A[r * n + j] * q->b[r];
        g_orig[j] = -v;
    }
    int total = 0, status = QPPVM_STATUS_OK;
    /* qpOASES solveRegularisedQP(): solve with H + eps I, then numRegularisationSteps proximal
     * re-solves with g <- g_orig - eps x_prev (only when the Hessian was regularised). */
    const int steps = (q->eps > 0.0) ? n_reg_steps : 0;
    for (int step = 0; step <= steps; ++step) {
        for (int j = 0; j < n; ++j) s->g[j] = g_orig[j] - (step > 0 ? q->eps * x[j] : 0.0);
        /* unconstrained minimiser xu = -J J^T g */
        double t[64];
        for (int j = 0; j < n; ++j) { double v = 0; for (int i = 0; i < n; ++i) v += s->J[i * n + j] * s->g[i]; t[j] = v; }
        for (int i = 0; i < n; ++i) { double v = 0; for (int j = 0; j < n; ++j) v += s->J[i * n + j] * t[j]; s->xu[i] = -v; }
        int it = 0;
        status = gi_solve(n, s->J, s->xu, q->nc, q->C, q->lA, q->uA, max_iter, y, &it, s->giwk, s->iwk, cand);
        if (cand) for (int i = 0; i < q->nc; ++i) cand[i] = (unsigned char)(cand[i] || y[i] != 0.0);   /* hot: the re-solve starts from this set */
        total += it;
        memcpy(x, s->xu, sizeof(double) * n);
        for (int j = 0; j < n && status == QPPVM_STATUS_OK; ++j) if (!isfinite(x[j])) status = QPPVM_STATUS_NUMERIC;
        if (status != QPPVM_STATUS_OK) break;
    }
    *iters = total;
    *kkt = (status == QPPVM_STATUS_OK) ? kkt_residual(q, s->g, x, y) : INFINITY;
    return status;
}

/* warm (optional, in/out): 8 words, the active rows of level 0 | level 1 of the previous solve of this problem
 * (what QPOases_sot carries from tick to tick: ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:64) -- tried first. */
int oracle_solve_record_warm(const qppvm_desc* d, const double* rec, void* out, double* diag, int mode, scratch* s, unsigned* warm);
int oracle_solve_record(const qppvm_desc* d, const double* rec, void* out, double* diag, int mode, scratch* s)
{
    return oracle_solve_record_warm(d, rec, out, diag, mode, s, NULL);
}
int oracle_solve_record_warm(const qppvm_desc* d, const double* rec, void* out, double* diag, int mode, scratch* s, unsigned* warm)
{
    qppvm_layout L;
    if (oracle_layout(d, &L)) return QPPVM_ERR_ARG;
    const int n = L.n_x, na = L.n_a, nv = L.n_v;
    double* xo = (double*)out;
    double* tau = xo + n;
    qppvm_trailer* tr = (qppvm_trailer*)(xo + n + na);
    memset(out, 0, L.out_bytes);
    if (diag) memset(diag, 0, sizeof(double) * L.diag_doubles);
    level_qp q; q.A = s->A; q.b = s->b; q.C = s->C; q.lA = s->lA; q.uA = s->uA;
    double x0[64], x1[64], y[128];
    int status = QPPVM_STATUS_OK, it0 = 0, it1 = 0;
    double kkt0 = INFINITY, kkt1 = INFINITY;
    for (int level = 0; level < 2 && status == QPPVM_STATUS_OK; ++level) {
        if (d->kind == QPPVM_KIND_FORCEACC) assemble_forceacc(d, &L, rec, level, x0, &q);
        else if (assemble_torque(d, &L, rec, level, x0, &q, s->wk + (160 + 64) * 64)) { status = QPPVM_STATUS_NUMERIC; break; }
        for (int i = 0; i < q.nc; ++i) if (!(isfinite(q.lA[i]) && isfinite(q.uA[i]))) status = QPPVM_STATUS_NUMERIC;
        if (status) break;
        double* x = level ? x1 : x0;
        unsigned char cand[128];
        if (warm) {
            for (int i = 0; i < q.nc; ++i) cand[i] = (unsigned char)(((warm[4 * level + (i >> 5)] >> (i & 31)) & 1u) || (level == 1 && i < L.row_opt && y[i] != 0.0));
        }
        status = solve_level(&q, mode, d->n_reg_steps, d->max_iter, s, x, y, level ? &it1 : &it0,
                             level ? &kkt1 : &kkt0, warm ? cand : NULL);
        if (warm && status == QPPVM_STATUS_OK) {
            for (int w = 0; w < 4; ++w) warm[4 * level + w] = 0u;
            for (int i = 0; i < q.nc; ++i) if (y[i] != 0.0 || q.lA[i] == q.uA[i]) warm[4 * level + (i >> 5)] |= 1u << (i & 31);
        }
        if (diag) {
            if (level == 0) memcpy(diag, x0, sizeof(double) * n);
            memcpy(diag + n + level * L.n_rows, y, sizeof(double) * q.nc);
            if (level == 1) for (int r = 0; r < QPPVM_M0; ++r) diag[n + 2 * L.n_rows + r] = q.lA[L.row_opt + r];
        }
        if (level == 1 && status == QPPVM_STATUS_OK)
            for (int i = 0; i < q.nc; ++i)
                if (y[i] != 0.0 || q.lA[i] == q.uA[i]) tr->active[i >> 5] |= 1u << (i & 31);
    }
    tr->status = status;
    tr->iters = (it0 & 0xffff) | (it1 << 16);
    tr->kkt[0] = (float)kkt0; tr->kkt[1] = (float)kkt1;
    const double* h = rec + L.off_h;
    if (status != QPPVM_STATUS_OK) {
        /* QPPVM: tau_qp = 0 then tau_d = 0 + h (ref:src/QPPVMPlugin.cpp:246-256);
         * ForceAcc: early return, nothing commanded (ref:src/ForceAcc.cpp:189-193) -> zeros. */
        memset(tr->active, 0, sizeof(tr->active));
        if (d->kind == QPPVM_KIND_TORQUE) for (int i = 0; i < na; ++i) tau[i] = h[i];
        return 0;
    }
    memcpy(xo, x1, sizeof(double) * n);
    if (d->kind == QPPVM_KIND_TORQUE) {
        for (int i = 0; i < na; ++i) tau[i] = x1[i] + h[i];
    } else {
        for (int a = 0; a < na; ++a) {                   /* tau = (M qdd + h - sum J_c^T [f;0])_actuated */
            double v = h[6 + a];
            for (int j = 0; j < nv; ++j) v += mget(rec + L.off_M, 6 + a, j) * x1[j];
            const int wd = (d->flags & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3;
            for (int ci = 0; ci < L.n_c; ++ci)
                for (int k = 0; k < wd; ++k) v -= rec[L.off_jc + (ci * 6 + k) * nv + 6 + a] * x1[nv + wd * ci + k];
            tau[a] = v;
        }
    }
    return 0;
}

/* Batch driver: `threads` <= 0 uses all cores.  Returns 0; per-problem status in the trailers. */
int oracle_solve_batch(const qppvm_desc* d, const double* recs, void* out, double* diag,
                       long long batch, int mode, int threads)
{
    qppvm_layout L;
    if (oracle_layout(d, &L)) return QPPVM_ERR_ARG;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    if ((long long)threads > batch) threads = batch > 0 ? (int)batch : 1;   /* no idle threads (and their scratch) on tiny batches */
#pragma omp parallel num_threads(threads)
    {
        scratch* s = scratch_new();
#pragma omp for schedule(dynamic, 16)
        for (long long i = 0; i < batch; ++i)
            oracle_solve_record(d, recs + i * L.rec_doubles, (char*)out + i * L.out_bytes,
                                diag ? diag + i * L.diag_doubles : NULL, mode, s);
        scratch_free(s);
    }
    return 0;
}

/* A tick sequence on ONE thread with the hot start a persistent solver gives: `warm` (8 words, in/out) carries the
 * working sets from record to record (zero it for a cold first tick).  The CPU side of the latency comparison. */
int oracle_solve_sequence(const qppvm_desc* d, const double* recs, void* out, long long ticks, int mode, unsigned* warm)
{
    qppvm_layout L;
    if (oracle_layout(d, &L) || !warm) return QPPVM_ERR_ARG;
    static __thread scratch* s = NULL;
    if (!s) s = scratch_new();
    for (long long i = 0; i < ticks; ++i)
        oracle_solve_record_warm(d, recs + i * (size_t)L.rec_doubles, (char*)out + i * (size_t)L.out_bytes, NULL, mode, s, warm);
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Generic dense QP entry (tests: analytic KATs, brute force, HiGHS cross-check).
 * min 1/2 ||A x - b||^2 + eps/2 ||x||^2  s.t. lA <= C x <= uA. */
int oracle_dense_qp(int n, int m, const double* A, const double* b, int nc, const double* C,
                    const double* lA, const double* uA, double eps, int n_reg_steps, int mode,
                    int max_iter, double* x, double* y, int* iters, double* kkt)
{
    if (n > 64 || m > 160 || nc > 128) return QPPVM_ERR_ARG;
    scratch* s = scratch_new();
    level_qp q; q.n = n; q.m = m; q.nc = nc; q.eps = eps;
    q.A = (double*)A; q.b = (double*)b; q.C = (double*)C; q.lA = (double*)lA; q.uA = (double*)uA;
    int st = solve_level(&q, mode, n_reg_steps, max_iter, s, x, y, iters, kkt, NULL);
    scratch_free(s);
    return st;
}

/* Explicit level matrices of one record (tests feed these to HiGHS / numpy KKT). */
int oracle_assemble(const qppvm_desc* d, const double* rec, int level, const double* x0,
                    double* A, double* b, double* C, double* lA, double* uA, int* dims /* m, nc */, double* eps)
{
    qppvm_layout L;
    if (oracle_layout(d, &L)) return QPPVM_ERR_ARG;
    level_qp q; q.A = A; q.b = b; q.C = C; q.lA = lA; q.uA = uA;
    if (d->kind == QPPVM_KIND_FORCEACC) assemble_forceacc(d, &L, rec, level, x0, &q);
    else {
        double* wk = (double*)malloc(sizeof(double) * 3 * 64 * 64);
        int e = assemble_torque(d, &L, rec, level, x0, &q, wk);
        free(wk);
        if (e) return QPPVM_ERR_ARG;
    }
    dims[0] = q.m; dims[1] = q.nc; *eps = q.eps;
    return 0;
}
