// rbd_kernel.cuh — on-device rigid-body front end: compact states -> QP records.
//
// What the reference obtains on the CPU each tick from XBot::ModelInterface after model->update()
// (ref:src/ForceAcc.cpp:256-282: sync_model; getJacobian :208; M, h inside DynamicFeasibility :109-114) plus the
// OpenSoT task right-hand sides (a_ref + lambda2 edot + lambda e) is produced here for a batch of
// states, straight into the record layout of include/qppvm_b200.h.  The CPU statement of the same arithmetic is
// qppvm_b200/gen.py: Robot.dynamics + records_from_states (tests compare the two to 1e-10).
//
// One CTA of 64 threads per state in flight, grid-stride over the batch.  The kinematic tree (<= 64 bodies) is a
// table in global memory; per-state body quantities live in shared memory:
//   1. forward kinematics by tree depth: R, p, joint axis z, angular velocity w, bias accelerations al / a
//   2. per body: COM offset, world inertia, bias wrench  m (a_c + g),  I al + w x I w
//   3. M(a,b) = sum_i m_i Jv_ia . Jv_ib + Jw_ia^T I_i Jw_ib  over the bodies i both columns move (ancestor masks),
//      h(a)   = sum_i Jv_ia . fv_i + Jw_ia . fw_i      -- one thread per packed entry, no atomics
//   4. task-link Jacobians, Jdot*qdot, task right-hand sides, bounds -> record
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qppvm_b200.h"

namespace qppvm {

constexpr int RBD_MAXB = 64;       // bodies
constexpr int RBD_TEAM = 64;

struct RobotTables {               // device pointers
    int n_a, n_b, max_depth;
    const int* parent;             // [n_b]
    const int* depth;              // [n_b]
    const unsigned long long* anc; // [n_b] bit k: joint k (body k+1) is on the path pelvis -> body
    const double* axis;            // [n_b][3]
    const double* offset;          // [n_b][3]
    const double* mass;            // [n_b]
    const double* com;             // [n_b][3]
    const double* inertia;         // [n_b][3]  diagonal, body frame
    const double* q_home;          // [n_a]
    const double* tau_max;         // [n_a]
    const int* contact_body;       // [4]
};

struct RbdShape {                  // record / state offsets (doubles), filled on the host from qppvm_layout
    int n_a, n_v, n_c, flags;
    int off_jwaist, off_jc, off_M, off_h, off_jdqd, off_rhs, off_taulim, off_cone, off_fbox, rec_doubles;
    int s_q, s_qd, s_R0, s_p0, s_tw, s_gains, s_ori, s_foot, s_mu, s_tscale, state_doubles;
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// Jacobian column `col` (generalised velocity index) of point `pt` on a body whose ancestor mask contains the
// column's joint: linear part jv, angular part jw.  Columns 0-2: base translation, 3-5: base rotation about p0.
__device__ __forceinline__ void jac_col(int col, const double* pt, const double* p0, const double* z, const double* p,
                                        double* jv, double* jw)
{
    if (col < 3) { jv[0] = col == 0; jv[1] = col == 1; jv[2] = col == 2; jw[0] = jw[1] = jw[2] = 0.0; return; }
    double ax[3], r[3];
    if (col < 6) {
        ax[0] = col == 3; ax[1] = col == 4; ax[2] = col == 5;
        r[0] = pt[0] - p0[0]; r[1] = pt[1] - p0[1]; r[2] = pt[2] - p0[2];
    } else {
        const int b = col - 5;                                // body moved by joint (col - 6)
        ax[0] = z[3 * b]; ax[1] = z[3 * b + 1]; ax[2] = z[3 * b + 2];
        r[0] = pt[0] - p[3 * b]; r[1] = pt[1] - p[3 * b + 1]; r[2] = pt[2] - p[3 * b + 2];
    }
    cross3(ax, r, jv);
    jw[0] = ax[0]; jw[1] = ax[1]; jw[2] = ax[2];
}

__device__ __forceinline__ bool col_moves(int col, unsigned long long anc)
{
    return col < 6 || ((anc >> (col - 6)) & 1ull);
}

__global__ void __launch_bounds__(RBD_TEAM)
rbd_records_kernel(RobotTables rob, RbdShape sh, const double* __restrict__ states, double* __restrict__ recs, long long batch)
{
    __shared__ double sR[RBD_MAXB * 9], sp[RBD_MAXB * 3], sz[RBD_MAXB * 3], sw[RBD_MAXB * 3], sal[RBD_MAXB * 3], sa[RBD_MAXB * 3];
    __shared__ double sc[RBD_MAXB * 3], sIw[RBD_MAXB * 9], sfv[RBD_MAXB * 3], sfw[RBD_MAXB * 3];
    __shared__ double sst[256];                     // the compact state
    __shared__ double sJ[6 * 64], sv[64];           // one task-link Jacobian, generalised velocity
    const int tid = threadIdx.x, nb = rob.n_b, nv = sh.n_v, na = sh.n_a, nc = sh.n_c;
    const double grav = 9.81;
    for (long long idx = blockIdx.x; idx < batch; idx += gridDim.x) {
        const double* st = states + idx * (size_t)sh.state_doubles;
        double* rec = recs + idx * (size_t)sh.rec_doubles;
        for (int i = tid; i < sh.state_doubles; i += RBD_TEAM) sst[i] = st[i];
        for (int i = tid; i < sh.rec_doubles; i += RBD_TEAM) rec[i] = 0.0;
        __syncthreads();
        const double* q = sst + sh.s_q; const double* qd = sst + sh.s_qd; const double* tw = sst + sh.s_tw;
        if (tid < nv) sv[tid] = tid < 6 ? tw[tid] : qd[tid - 6];
        if (tid == 0) {                             // body 0 = floating base
            for (int k = 0; k < 9; ++k) sR[k] = sst[sh.s_R0 + k];
            for (int k = 0; k < 3; ++k) { sp[k] = sst[sh.s_p0 + k]; sz[k] = 0.0; sw[k] = tw[3 + k]; sal[k] = 0.0; sa[k] = 0.0; }
        }
        __syncthreads();
        // ---- 1. forward kinematics, velocities and bias accelerations, one tree level at a time
        for (int d = 1; d <= rob.max_depth; ++d) {
            for (int i = tid; i < nb; i += RBD_TEAM) {
                if (rob.depth[i] != d) continue;
                const int pa = rob.parent[i];
                const double* Rp = sR + 9 * pa;
                const double ax[3] = {rob.axis[3 * i], rob.axis[3 * i + 1], rob.axis[3 * i + 2]};
                const double of[3] = {rob.offset[3 * i], rob.offset[3 * i + 1], rob.offset[3 * i + 2]};
                double r[3], zz[3];
                for (int a = 0; a < 3; ++a) {
                    r[a] = Rp[3 * a] * of[0] + Rp[3 * a + 1] * of[1] + Rp[3 * a + 2] * of[2];
                    zz[a] = Rp[3 * a] * ax[0] + Rp[3 * a + 1] * ax[1] + Rp[3 * a + 2] * ax[2];
                    sp[3 * i + a] = sp[3 * pa + a] + r[a];
                    sz[3 * i + a] = zz[a];
                }
                // Rodrigues about the unit joint axis: Rot = I + s K + (1 - c) K^2
                double sn, cs;
                sincos(q[i - 1], &sn, &cs);
                const double K[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
                double Rot[9];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) {
                        const double kk = K[3 * a] * K[b] + K[3 * a + 1] * K[3 + b] + K[3 * a + 2] * K[6 + b];
                        Rot[3 * a + b] = (a == b ? 1.0 : 0.0) + sn * K[3 * a + b] + (1.0 - cs) * kk;
                    }
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        sR[9 * i + 3 * a + b] = Rp[3 * a] * Rot[b] + Rp[3 * a + 1] * Rot[3 + b] + Rp[3 * a + 2] * Rot[6 + b];
                const double qdi = qd[i - 1];
                const double zq[3] = {zz[0] * qdi, zz[1] * qdi, zz[2] * qdi};
                const double* wp = sw + 3 * pa; const double* alp = sal + 3 * pa;
                double c1[3], c2[3], c3[3];
                cross3(wp, zq, c1);
                cross3(alp, r, c2);
                cross3(wp, r, c3);
                double c4[3];
                cross3(wp, c3, c4);
                for (int a = 0; a < 3; ++a) {
                    sw[3 * i + a] = wp[a] + zq[a];
                    sal[3 * i + a] = alp[a] + c1[a];
                    sa[3 * i + a] = sa[3 * pa + a] + c2[a] + c4[a];
                }
            }
            __syncthreads();
        }
        // ---- 2. per body: COM offset, world inertia, bias wrench
        for (int i = tid; i < nb; i += RBD_TEAM) {
            const double* R = sR + 9 * i;
            const double cm[3] = {rob.com[3 * i], rob.com[3 * i + 1], rob.com[3 * i + 2]};
            const double In[3] = {rob.inertia[3 * i], rob.inertia[3 * i + 1], rob.inertia[3 * i + 2]};
            const double m = rob.mass[i];
            double c[3];
            for (int a = 0; a < 3; ++a) c[a] = R[3 * a] * cm[0] + R[3 * a + 1] * cm[1] + R[3 * a + 2] * cm[2];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    sIw[9 * i + 3 * a + b] = R[3 * a] * In[0] * R[3 * b] + R[3 * a + 1] * In[1] * R[3 * b + 1] + R[3 * a + 2] * In[2] * R[3 * b + 2];
            const double* w = sw + 3 * i; const double* al = sal + 3 * i;
            double t1[3], t2[3], t3[3];
            cross3(al, c, t1);
            cross3(w, c, t2);
            cross3(w, t2, t3);
            double Iw_w[3], Iw_al[3], t4[3];
            for (int a = 0; a < 3; ++a) {
                Iw_w[a] = sIw[9 * i + 3 * a] * w[0] + sIw[9 * i + 3 * a + 1] * w[1] + sIw[9 * i + 3 * a + 2] * w[2];
                Iw_al[a] = sIw[9 * i + 3 * a] * al[0] + sIw[9 * i + 3 * a + 1] * al[1] + sIw[9 * i + 3 * a + 2] * al[2];
            }
            cross3(w, Iw_w, t4);
            for (int a = 0; a < 3; ++a) {
                sc[3 * i + a] = sp[3 * i + a] + c[a];         // COM position in the world
                sfv[3 * i + a] = m * (sa[3 * i + a] + t1[a] + t3[a] + (a == 2 ? grav : 0.0));
                sfw[3 * i + a] = Iw_al[a] + t4[a];
            }
        }
        __syncthreads();
        // ---- 3. mass matrix (packed lower) and nonlinear term
        const int nM = nv * (nv + 1) / 2;
        for (int e = tid; e < nM + nv; e += RBD_TEAM) {
            if (e < nM) {
                int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while (a * (a + 1) / 2 > e) --a;
                while ((a + 1) * (a + 2) / 2 <= e) ++a;
                const int b = e - a * (a + 1) / 2;            // a >= b
                double acc = 0.0;
                for (int i = 0; i < nb; ++i) {
                    const unsigned long long an = rob.anc[i];
                    if (!col_moves(a, an) || !col_moves(b, an)) continue;
                    double jva[3], jwa[3], jvb[3], jwb[3];
                    jac_col(a, sc + 3 * i, sp, sz, sp, jva, jwa);
                    jac_col(b, sc + 3 * i, sp, sz, sp, jvb, jwb);
                    const double* I = sIw + 9 * i;
                    const double Ib[3] = {I[0] * jwb[0] + I[1] * jwb[1] + I[2] * jwb[2], I[3] * jwb[0] + I[4] * jwb[1] + I[5] * jwb[2],
                                          I[6] * jwb[0] + I[7] * jwb[1] + I[8] * jwb[2]};
                    acc += rob.mass[i] * (jva[0] * jvb[0] + jva[1] * jvb[1] + jva[2] * jvb[2]) + (jwa[0] * Ib[0] + jwa[1] * Ib[1] + jwa[2] * Ib[2]);
                }
                rec[sh.off_M + e] = acc;
            } else {
                const int a = e - nM;
                double acc = 0.0;
                for (int i = 0; i < nb; ++i) {
                    if (!col_moves(a, rob.anc[i])) continue;
                    double jv[3], jw[3];
                    jac_col(a, sc + 3 * i, sp, sz, sp, jv, jw);
                    acc += jv[0] * sfv[3 * i] + jv[1] * sfv[3 * i + 1] + jv[2] * sfv[3 * i + 2]
                         + jw[0] * sfw[3 * i] + jw[1] * sfw[3 * i + 1] + jw[2] * sfw[3 * i + 2];
                }
                rec[sh.off_h + a] = acc;
            }
        }
        // ---- 4. task links: waist (body 0) then the contact links
        const double* gains = sst + sh.s_gains;
        const double lam_w = 100.0 * gains[0], lam2_w = 20.0 * gains[1], lam_p = 100.0 * gains[2], lam2_p = 20.0 * gains[3];
        for (int t = 0; t <= nc; ++t) {
            const int body = t == 0 ? 0 : rob.contact_body[t - 1];
            const unsigned long long an = rob.anc[body];
            __syncthreads();
            for (int col = tid; col < nv; col += RBD_TEAM) {
                double jv[3] = {0, 0, 0}, jw[3] = {0, 0, 0};
                if (col_moves(col, an)) jac_col(col, sp + 3 * body, sp, sz, sp, jv, jw);
                for (int r = 0; r < 3; ++r) { sJ[r * 64 + col] = jv[r]; sJ[(3 + r) * 64 + col] = jw[r]; }
                double* Jout = rec + (t == 0 ? sh.off_jwaist : sh.off_jc + (t - 1) * 6 * nv);
                for (int r = 0; r < 3; ++r) { Jout[r * nv + col] = jv[r]; Jout[(3 + r) * nv + col] = jw[r]; }
            }
            __syncthreads();
            if (tid < 6) {
                double jvel = 0.0;
                for (int col = 0; col < nv; ++col) jvel += sJ[tid * 64 + col] * sv[col];
                rec[sh.off_jdqd + 6 * t + tid] = tid < 3 ? sa[3 * body + tid] : sal[3 * body + tid - 3];
                double e;
                if (t == 0) e = tid < 3 ? (tid == 2 ? -0.1 : 0.0) : sst[sh.s_ori + tid - 3];      // ref:src/ForceAcc.cpp:181
                else e = sst[sh.s_foot + 6 * (t - 1) + tid];
                rec[sh.off_rhs + 6 * t + tid] = (t == 0 ? lam_w : lam_p) * e - (t == 0 ? lam2_w : lam2_p) * jvel;
            }
            if (t > 0 && tid < 10 && (sh.flags & QPPVM_FLAG_FRICTION_CONES))
                rec[sh.off_cone + 10 * (t - 1) + tid] = tid < 9 ? sR[9 * body + tid] : sst[sh.s_mu + t - 1];
            if (t > 0 && tid < 6) {
                const double fb[6] = {-1000.0, -1000.0, 10.0, 1000.0, 1000.0, 1000.0};              // ref:src/ForceAcc.cpp:75-76
                rec[sh.off_fbox + 6 * (t - 1) + tid] = fb[tid];
            }
        }
        // postural right-hand side and torque limits
        for (int j = tid; j < nv; j += RBD_TEAM) {
            const double e = j < 6 ? 0.0 : rob.q_home[j - 6] - q[j - 6];
            rec[sh.off_rhs + 6 * (1 + nc) + j] = lam_p * e - lam2_p * sv[j];
        }
        if (sh.flags & QPPVM_FLAG_TORQUE_LIMITS)
            for (int a = tid; a < na; a += RBD_TEAM) {
                const double tm = rob.tau_max[a] * sst[sh.s_tscale + a];
                rec[sh.off_taulim + a] = -tm;
                rec[sh.off_taulim + na + a] = tm;
            }
        __syncthreads();
    }
}

}  // namespace qppvm
