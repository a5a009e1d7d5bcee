// rbd_kernel.cuh — on-device rigid-body front end: compact states -> QP records.
//
// What the reference obtains on the CPU each tick from XBot::ModelInterface after model->update()
// (ref:src/ForceAcc.cpp:256-282: sync_model; getJacobian :208; M, h inside DynamicFeasibility :109-114) plus the
// OpenSoT task right-hand sides (a_ref + lambda2 edot + lambda e) is produced here for a batch of
// states, straight into the record layout of include/qppvm_b200.h.  The CPU statement of the same arithmetic is
// qppvm_b200/gen.py: Robot.dynamics + records_from_states (tests compare the two to 1e-10).
//
// One thread per state.  The kinematic tree (<= 64 bodies) is a table staged in shared memory; per-state body
// quantities live in per-thread local memory:
//   1. forward kinematics by tree depth: R, p, joint axis z, angular velocity w, bias accelerations al / a
//   2. per body: COM offset, world inertia, bias wrench  m (a_c + g),  I al + w x I w
//   3. composite-rigid-body form of  M(a,b) = sum_i m_i Jv_ia . Jv_ib + Jw_ia^T I_i Jw_ib  and
//      h(a) = sum_i Jv_ia . fv_i + Jw_ia . fw_i :  spatial inertia (m, m c, I about the world origin) and bias wrench
//      are accumulated over subtrees leaf -> root in a fixed order (no atomics), each generalised-velocity column
//      gets its twist about the world origin and the momentum of the subtree it moves, and one thread per packed
//      entry takes a 6-term dot product
//   4. task-link Jacobians, Jdot*qdot, task right-hand sides, bounds -> record
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qppvm_b200.h"

namespace qppvm {

constexpr int RBD_MAXB = 40;       // bodies (n_a <= 39)
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
    int s_q, s_qd, s_R0, s_p0, s_tw, s_gains, s_ori, s_foot, s_mu, s_tscale, s_wpos, state_doubles;
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// One THREAD per state (every lane busy; a CTA-per-state mapping left <= 5 of 64 threads active per tree level
// and was instruction-bound: profiles/README.md).  Per-body quantities live in per-thread local memory, which the
// hardware interleaves across the threads of a warp, so the body loops are coalesced; the kinematic-tree tables
// are staged once per CTA in shared memory and read as broadcasts.
__global__ void __launch_bounds__(RBD_TEAM)
rbd_records_kernel(RobotTables rob, RbdShape sh, const double* __restrict__ states, double* __restrict__ recs, long long batch)
{
    __shared__ double tAxis[RBD_MAXB * 3], tOff[RBD_MAXB * 3], tCom[RBD_MAXB * 3], tIn[RBD_MAXB * 3], tMass[RBD_MAXB];
    __shared__ double tQh[RBD_MAXB], tTm[RBD_MAXB];
    __shared__ unsigned long long tAnc[RBD_MAXB];
    __shared__ int tParent[RBD_MAXB], tContact[4];
    const int tid = threadIdx.x, nb = rob.n_b, nv = sh.n_v, na = sh.n_a, nc = sh.n_c;
    const double grav = 9.81;
    for (int i = tid; i < nb * 3; i += RBD_TEAM) { tAxis[i] = rob.axis[i]; tOff[i] = rob.offset[i]; tCom[i] = rob.com[i]; tIn[i] = rob.inertia[i]; }
    for (int i = tid; i < nb; i += RBD_TEAM) { tMass[i] = rob.mass[i]; tAnc[i] = rob.anc[i]; tParent[i] = rob.parent[i]; }
    for (int i = tid; i < na; i += RBD_TEAM) { tQh[i] = rob.q_home[i]; tTm[i] = rob.tau_max[i]; }
    if (tid < 4) tContact[tid] = rob.contact_body[tid];
    __syncthreads();
    const long long idx = (long long)blockIdx.x * RBD_TEAM + tid;
    if (idx >= batch) return;
    const double* st = states + idx * (size_t)sh.state_doubles;
    double* rec = recs + idx * (size_t)sh.rec_doubles;
    const double* q = st + sh.s_q; const double* qd = st + sh.s_qd; const double* tw = st + sh.s_tw;

    double bR[RBD_MAXB][9], bp[RBD_MAXB][3], bz[RBD_MAXB][3], bw[RBD_MAXB][3], bal[RBD_MAXB][3], ba[RBD_MAXB][3];
    double bC[RBD_MAXB][16];
    for (int k = 0; k < 9; ++k) bR[0][k] = st[sh.s_R0 + k];
    for (int k = 0; k < 3; ++k) { bp[0][k] = st[sh.s_p0 + k]; bz[0][k] = 0.0; bw[0][k] = tw[3 + k]; bal[0][k] = 0.0; ba[0][k] = 0.0; }
    // ---- 1. forward kinematics, velocities and bias accelerations (bodies are topologically ordered)
#pragma unroll 1
    for (int i = 1; i < nb; ++i) {
        const int pa = tParent[i];
        double Rp[9], ax[3], of[3], r[3], zz[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rp[k] = bR[pa][k];
#pragma unroll
        for (int k = 0; k < 3; ++k) { ax[k] = tAxis[3 * i + k]; of[k] = tOff[3 * i + k]; }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            r[a] = Rp[3 * a] * of[0] + Rp[3 * a + 1] * of[1] + Rp[3 * a + 2] * of[2];
            zz[a] = Rp[3 * a] * ax[0] + Rp[3 * a + 1] * ax[1] + Rp[3 * a + 2] * ax[2];
            bp[i][a] = bp[pa][a] + r[a];
            bz[i][a] = zz[a];
        }
        double sn, cs;                              // Rodrigues about the unit joint axis: Rot = I + s K + (1 - c) K^2
        sincos(q[i - 1], &sn, &cs);
        const double K[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
        double Rot[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const double kk = K[3 * a] * K[b] + K[3 * a + 1] * K[3 + b] + K[3 * a + 2] * K[6 + b];
                Rot[3 * a + b] = (a == b ? 1.0 : 0.0) + sn * K[3 * a + b] + (1.0 - cs) * kk;
            }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                bR[i][3 * a + b] = Rp[3 * a] * Rot[b] + Rp[3 * a + 1] * Rot[3 + b] + Rp[3 * a + 2] * Rot[6 + b];
        const double qdi = qd[i - 1];
        const double zq[3] = {zz[0] * qdi, zz[1] * qdi, zz[2] * qdi};
        double wp[3], alp[3], c1[3], c2[3], c3[3], c4[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { wp[k] = bw[pa][k]; alp[k] = bal[pa][k]; }
        cross3(wp, zq, c1);
        cross3(alp, r, c2);
        cross3(wp, r, c3);
        cross3(wp, c3, c4);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            bw[i][a] = wp[a] + zq[a];
            bal[i][a] = alp[a] + c1[a];
            ba[i][a] = ba[pa][a] + c2[a] + c4[a];
        }
    }
    // ---- 2. per body: spatial inertia about the WORLD ORIGIN (m | m c | I_O, 10 numbers) and bias wrench about
    // the origin (force | moment, 6 numbers)
#pragma unroll 1
    for (int i = 0; i < nb; ++i) {
        double R[9], w[3], al[3], c[3], Iw[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = bR[i][k];
#pragma unroll
        for (int k = 0; k < 3; ++k) { w[k] = bw[i][k]; al[k] = bal[i][k]; }
        const double cm[3] = {tCom[3 * i], tCom[3 * i + 1], tCom[3 * i + 2]};
        const double In[3] = {tIn[3 * i], tIn[3 * i + 1], tIn[3 * i + 2]};
        const double m = tMass[i];
#pragma unroll
        for (int a = 0; a < 3; ++a) c[a] = R[3 * a] * cm[0] + R[3 * a + 1] * cm[1] + R[3 * a + 2] * cm[2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                Iw[3 * a + b] = R[3 * a] * In[0] * R[3 * b] + R[3 * a + 1] * In[1] * R[3 * b + 1] + R[3 * a + 2] * In[2] * R[3 * b + 2];
        double t1[3], t2[3], t3[3], Iw_w[3], Iw_al[3], t4[3], fv[3], fw[3], pc[3], mom[3];
        cross3(al, c, t1);
        cross3(w, c, t2);
        cross3(w, t2, t3);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            Iw_w[a] = Iw[3 * a] * w[0] + Iw[3 * a + 1] * w[1] + Iw[3 * a + 2] * w[2];
            Iw_al[a] = Iw[3 * a] * al[0] + Iw[3 * a + 1] * al[1] + Iw[3 * a + 2] * al[2];
        }
        cross3(w, Iw_w, t4);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            pc[a] = bp[i][a] + c[a];                          // COM position in the world
            fv[a] = m * (ba[i][a] + t1[a] + t3[a] + (a == 2 ? grav : 0.0));
            fw[a] = Iw_al[a] + t4[a];
        }
        cross3(pc, fv, mom);
        const double cc = pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2];
        bC[i][0] = m; bC[i][1] = m * pc[0]; bC[i][2] = m * pc[1]; bC[i][3] = m * pc[2];
        // I_O = Iw + m (|c|^2 1 - c c^T), symmetric: xx xy xz yy yz zz
        bC[i][4] = Iw[0] + m * (cc - pc[0] * pc[0]); bC[i][5] = Iw[1] - m * pc[0] * pc[1]; bC[i][6] = Iw[2] - m * pc[0] * pc[2];
        bC[i][7] = Iw[4] + m * (cc - pc[1] * pc[1]); bC[i][8] = Iw[5] - m * pc[1] * pc[2]; bC[i][9] = Iw[8] + m * (cc - pc[2] * pc[2]);
        bC[i][10] = fv[0]; bC[i][11] = fv[1]; bC[i][12] = fv[2];
        bC[i][13] = fw[0] + mom[0]; bC[i][14] = fw[1] + mom[1]; bC[i][15] = fw[2] + mom[2];
    }
    // composites: one backward sweep folds every body into its parent (parent < child), fixed order
#pragma unroll 1
    for (int j = nb - 1; j >= 1; --j) {
        const int pa = tParent[j];
        double cj[16], cp[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { cj[k] = bC[j][k]; cp[k] = bC[pa][k]; }       // 32 loads in flight, then 16 stores
#pragma unroll
        for (int k = 0; k < 16; ++k) bC[pa][k] = cp[k] + cj[k];
    }
    // ---- 3. per column: twist about the world origin (omega | v_O) and the momentum (n | l) of the subtree it moves;
    // M(a,b) = twist_b . momentum_a for related columns (a >= b, column a the deeper one)
    double T[RBD_MAXB + 6][12];
#pragma unroll 1
    for (int col = 0; col < nv; ++col) {
        double om[3] = {0, 0, 0}, vo[3] = {0, 0, 0}, p0[3] = {bp[0][0], bp[0][1], bp[0][2]};
        int body = 0;                                         // base columns move the whole robot
        if (col < 3) vo[col] = 1.0;
        else if (col < 6) { om[col - 3] = 1.0; cross3(p0, om, vo); }                 // v_O = p0 x e_k
        else {
            body = col - 5;
            double pb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { om[k] = bz[body][k]; pb[k] = bp[body][k]; }
            cross3(pb, om, vo);
        }
        double C[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) C[k] = bC[body][k];
        const double mc[3] = {C[1], C[2], C[3]};
        double l[3], n[3], t[3];
        cross3(om, mc, t);                                    // linear momentum  l = m v_O + omega x (m c)
#pragma unroll
        for (int a = 0; a < 3; ++a) l[a] = C[0] * vo[a] + t[a];
        cross3(mc, vo, t);                                    // angular momentum n = (m c) x v_O + I_O omega
        n[0] = t[0] + C[4] * om[0] + C[5] * om[1] + C[6] * om[2];
        n[1] = t[1] + C[5] * om[0] + C[7] * om[1] + C[8] * om[2];
        n[2] = t[2] + C[6] * om[0] + C[8] * om[1] + C[9] * om[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) { T[col][a] = om[a]; T[col][3 + a] = vo[a]; T[col][6 + a] = n[a]; T[col][9 + a] = l[a]; }
        rec[sh.off_h + col] = om[0] * C[13] + om[1] * C[14] + om[2] * C[15] + vo[0] * C[10] + vo[1] * C[11] + vo[2] * C[12];
    }
#pragma unroll 1
    for (int a = 0; a < nv; ++a) {
        double Ta[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) Ta[k] = T[a][6 + k];
        const unsigned long long anc_a = a >= 6 ? tAnc[a - 5] : 0ull;
        // (the twists are loaded unconditionally and four columns at a time: the per-thread arrays live in local memory,
        // i.e. in L2 / DRAM at this footprint, and the kernel is bound by how many of those loads are in flight --
        // profiles/README.md, round 2: 94 % of the stall samples were long-scoreboard waits with one column per iteration)
        double* const Mrow = rec + sh.off_M + a * (a + 1) / 2;
#pragma unroll 4
        for (int b = 0; b <= a; ++b) {
            const bool related = b < 6 || a == b || ((anc_a >> (b < 6 ? 0 : b - 6)) & 1ull);     // joint b on the path to body of a
            const double val = T[b][0] * Ta[0] + T[b][1] * Ta[1] + T[b][2] * Ta[2] + T[b][3] * Ta[3] + T[b][4] * Ta[4] + T[b][5] * Ta[5];
            Mrow[b] = related ? val : 0.0;
        }
    }
    // ---- 4. task links: waist (body 0) then the contact links
    const double* gains = st + sh.s_gains;
    const double lam_w = 100.0 * gains[0], lam2_w = 20.0 * gains[1], lam_p = 100.0 * gains[2], lam2_p = 20.0 * gains[3];
#pragma unroll 1
    for (int t = 0; t <= nc; ++t) {
        const int body = t == 0 ? 0 : tContact[t - 1];
        const unsigned long long an = tAnc[body];
        double pt[3] = {bp[body][0], bp[body][1], bp[body][2]}, p0[3] = {bp[0][0], bp[0][1], bp[0][2]};
        double jvel[6] = {0, 0, 0, 0, 0, 0};
        double* Jout = rec + (t == 0 ? sh.off_jwaist : sh.off_jc + (t - 1) * 6 * nv);
#pragma unroll
        for (int col = 0; col < 6; ++col) {                    // floating-base columns
            double jv[3] = {0, 0, 0}, jw[3] = {0, 0, 0};
            if (col < 3) jv[col] = 1.0;
            else { double r[3] = {pt[0] - p0[0], pt[1] - p0[1], pt[2] - p0[2]}; jw[col - 3] = 1.0; cross3(jw, r, jv); }
            const double vc = tw[col];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Jout[r * nv + col] = jv[r]; Jout[(3 + r) * nv + col] = jw[r];
                jvel[r] = fma(jv[r], vc, jvel[r]); jvel[3 + r] = fma(jw[r], vc, jvel[3 + r]);
            }
        }
        // joint columns: axis and origin of every joint are fetched whether or not it moves this link (loads of four
        // columns in flight), the ancestor mask selects afterwards
#pragma unroll 4
        for (int col = 6; col < nv; ++col) {
            const int b = col - 5;
            const bool on = (an >> (col - 6)) & 1ull;
            double jv[3], jw[3], r[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { jw[k] = bz[b][k]; r[k] = pt[k] - bp[b][k]; }
            cross3(jw, r, jv);
            const double vc = qd[col - 6];
#pragma unroll
            for (int r2 = 0; r2 < 3; ++r2) {
                const double v = on ? jv[r2] : 0.0, w = on ? jw[r2] : 0.0;
                Jout[r2 * nv + col] = v; Jout[(3 + r2) * nv + col] = w;
                jvel[r2] = fma(v, vc, jvel[r2]); jvel[3 + r2] = fma(w, vc, jvel[3 + r2]);
            }
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            rec[sh.off_jdqd + 6 * t + r] = r < 3 ? ba[body][r] : bal[body][r - 3];
            double e;
            if (t == 0) e = r < 3 ? st[sh.s_wpos + r] : st[sh.s_ori + r - 3];   // reference = initial - 0.1 z (ref:src/ForceAcc.cpp:181)
            else e = st[sh.s_foot + 6 * (t - 1) + r];
            rec[sh.off_rhs + 6 * t + r] = (t == 0 ? lam_w : lam_p) * e - (t == 0 ? lam2_w : lam2_p) * jvel[r];
        }
        if (t > 0) {
            if (sh.flags & QPPVM_FLAG_FRICTION_CONES) {
                for (int k = 0; k < 9; ++k) rec[sh.off_cone + 10 * (t - 1) + k] = bR[body][k];
                rec[sh.off_cone + 10 * (t - 1) + 9] = st[sh.s_mu + t - 1];
            }
            const double fb[6] = {-1000.0, -1000.0, 10.0, 1000.0, 1000.0, 1000.0};                  // ref:src/ForceAcc.cpp:75-76
#pragma unroll
            for (int k = 0; k < 6; ++k) rec[sh.off_fbox + 6 * (t - 1) + k] = fb[k];
        }
    }
    // postural right-hand side and torque limits
#pragma unroll 1
    for (int j = 0; j < nv; ++j) {
        const double e = j < 6 ? 0.0 : tQh[j - 6] - q[j - 6];
        const double vc = j < 6 ? tw[j] : qd[j - 6];
        rec[sh.off_rhs + 6 * (1 + nc) + j] = lam_p * e - lam2_p * vc;
    }
    if (sh.flags & QPPVM_FLAG_TORQUE_LIMITS)
#pragma unroll 1
        for (int a = 0; a < na; ++a) {
            const double tm = tTm[a] * st[sh.s_tscale + a];
            rec[sh.off_taulim + a] = -tm;
            rec[sh.off_taulim + na + a] = tm;
        }
    if (sh.rec_doubles & 1 || true) {                         // padding double of the record stride
        for (int k = sh.off_fbox + 6 * nc; k < sh.rec_doubles; ++k) rec[k] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// Warp-per-state front end (round 2).  The thread-per-state kernel above keeps 17 KB of per-body data per thread in
// local memory: at any useful batch that footprint lives in L2 / DRAM (ncu: 5.6 GB of DRAM traffic for 1.2 GB of
// records, 94 % of the stall samples on local loads, issue slots 4 % busy), and because every state is one serial
// thread a chunk of a few thousand states takes as long as 65 536 (2.6 ms: one wave).  Here one WARP owns a state and
// the per-body data sits in shared memory (14.6 KB per state for the 33-DoF robot): lanes take the bodies of one tree
// level (forward kinematics), bodies (inertias), columns (twists / momenta), packed entries of M, Jacobian columns;
// every store to the record is a coalesced 256-byte row of lanes.  Same arithmetic as above and as gen.py.
// Per-state block (doubles): state copy | bodies x RBD_BS [R 9 | p 3 | z 3 | w 3 | al 3 | a 3 | C 16] | columns x RBD_TS
// [omega 3 | v_O 3 | n 3 | l 3].  Odd strides: a lane per body / column walks conflict-free banks.
// ------------------------------------------------------------------------------------------
constexpr int RBD_WARPS = 4;       // states in flight per CTA
constexpr int RBD_BS = 41, RBD_TS = 13;
__host__ __device__ inline int rbd_state_pad(const RbdShape& sh) { return (sh.state_doubles + 1) & ~1; }
__host__ __device__ inline int rbd_warp_doubles(const RbdShape& sh) { return rbd_state_pad(sh) + RBD_BS * (sh.n_a + 1) + RBD_TS * sh.n_v + 1; }

__global__ void __launch_bounds__(32 * RBD_WARPS)
rbd_records_warp_kernel(RobotTables rob, RbdShape sh, const double* __restrict__ states, double* __restrict__ recs, long long batch)
{
    __shared__ double tAxis[RBD_MAXB * 3], tOff[RBD_MAXB * 3], tCom[RBD_MAXB * 3], tIn[RBD_MAXB * 3], tMass[RBD_MAXB];
    __shared__ double tQh[RBD_MAXB], tTm[RBD_MAXB];
    __shared__ unsigned long long tAnc[RBD_MAXB];
    __shared__ int tParent[RBD_MAXB], tContact[4], tOrder[RBD_MAXB], tLevel[RBD_MAXB + 2];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    const int nb = rob.n_b, nv = sh.n_v, na = sh.n_a, nc = sh.n_c, maxd = rob.max_depth;
    const double grav = 9.81;
    for (int i = tid; i < nb * 3; i += 32 * RBD_WARPS) { tAxis[i] = rob.axis[i]; tOff[i] = rob.offset[i]; tCom[i] = rob.com[i]; tIn[i] = rob.inertia[i]; }
    for (int i = tid; i < nb; i += 32 * RBD_WARPS) { tMass[i] = rob.mass[i]; tAnc[i] = rob.anc[i]; tParent[i] = rob.parent[i]; }
    for (int i = tid; i < na; i += 32 * RBD_WARPS) { tQh[i] = rob.q_home[i]; tTm[i] = rob.tau_max[i]; }
    if (tid < 4) tContact[tid] = rob.contact_body[tid];
    // bodies sorted by tree depth (stable): tOrder, level d occupies [tLevel[d], tLevel[d + 1])
    for (int i = tid; i < nb; i += 32 * RBD_WARPS) {
        const int di = rob.depth[i];
        int rank = 0;
        for (int j = 0; j < nb; ++j) { const int dj = rob.depth[j]; rank += (dj < di || (dj == di && j < i)) ? 1 : 0; }
        tOrder[rank] = i;
    }
    for (int d = tid; d <= maxd + 1; d += 32 * RBD_WARPS) {
        int cnt = 0;
        for (int j = 0; j < nb; ++j) cnt += rob.depth[j] < d ? 1 : 0;
        tLevel[d] = cnt;
    }
    __syncthreads();
    const int SDP = rbd_state_pad(sh);
    double* const S = reinterpret_cast<double*>(g_smem) + (size_t)wp * rbd_warp_doubles(sh);
    double* const Bd = S + SDP;
    double* const T = Bd + RBD_BS * nb;
    const double* const q = S + sh.s_q; const double* const qd = S + sh.s_qd; const double* const tw = S + sh.s_tw;
#pragma unroll 1
    for (long long idx = (long long)blockIdx.x * RBD_WARPS + wp; idx < batch; idx += (long long)gridDim.x * RBD_WARPS) {
        const double* st = states + idx * (size_t)sh.state_doubles;
        double* rec = recs + idx * (size_t)sh.rec_doubles;
        __syncwarp();                                         // the previous state's readers are done with S / Bd / T
        for (int k = lane; k < sh.state_doubles; k += 32) S[k] = st[k];
        __syncwarp();
        // body 0: the floating base
        if (lane < 9) Bd[lane] = S[sh.s_R0 + lane];
        else if (lane < 12) Bd[lane] = S[sh.s_p0 + lane - 9];
        else if (lane < 15) Bd[lane] = 0.0;                   // z
        else if (lane < 18) Bd[lane] = tw[3 + lane - 15];     // w
        else if (lane < 24) Bd[lane] = 0.0;                   // al, a
        __syncwarp();
        // ---- 1. forward kinematics, velocities and bias accelerations: one tree level at a time, a lane per body
#pragma unroll 1
        for (int d = 1; d <= maxd; ++d) {
#pragma unroll 1
            for (int kk = tLevel[d] + lane; kk < tLevel[d + 1]; kk += 32) {
                const int i = tOrder[kk], pa = tParent[i];
                const double* P = Bd + RBD_BS * pa;
                double* Bi = Bd + RBD_BS * i;
                double Rp[9], ax[3], of[3], r[3], zz[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) Rp[k] = P[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) { ax[k] = tAxis[3 * i + k]; of[k] = tOff[3 * i + k]; }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    r[a] = Rp[3 * a] * of[0] + Rp[3 * a + 1] * of[1] + Rp[3 * a + 2] * of[2];
                    zz[a] = Rp[3 * a] * ax[0] + Rp[3 * a + 1] * ax[1] + Rp[3 * a + 2] * ax[2];
                    Bi[9 + a] = P[9 + a] + r[a];
                    Bi[12 + a] = zz[a];
                }
                double sn, cs;                              // Rodrigues about the unit joint axis: Rot = I + s K + (1 - c) K^2
                sincos(q[i - 1], &sn, &cs);
                const double K[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
                double Rot[9];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const double kq = K[3 * a] * K[b] + K[3 * a + 1] * K[3 + b] + K[3 * a + 2] * K[6 + b];
                        Rot[3 * a + b] = (a == b ? 1.0 : 0.0) + sn * K[3 * a + b] + (1.0 - cs) * kq;
                    }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b)
                        Bi[3 * a + b] = Rp[3 * a] * Rot[b] + Rp[3 * a + 1] * Rot[3 + b] + Rp[3 * a + 2] * Rot[6 + b];
                const double qdi = qd[i - 1];
                const double zq[3] = {zz[0] * qdi, zz[1] * qdi, zz[2] * qdi};
                double wpar[3], alp[3], c1[3], c2[3], c3[3], c4[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) { wpar[k] = P[15 + k]; alp[k] = P[18 + k]; }
                cross3(wpar, zq, c1);
                cross3(alp, r, c2);
                cross3(wpar, r, c3);
                cross3(wpar, c3, c4);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    Bi[15 + a] = wpar[a] + zq[a];
                    Bi[18 + a] = alp[a] + c1[a];
                    Bi[21 + a] = P[21 + a] + c2[a] + c4[a];
                }
            }
            __syncwarp();
        }
        // ---- 2. per body: spatial inertia about the world origin (m | m c | I_O) and bias wrench about the origin
#pragma unroll 1
        for (int i = lane; i < nb; i += 32) {
            double* Bi = Bd + RBD_BS * i;
            double R[9], w[3], al[3], c[3], Iw[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Bi[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) { w[k] = Bi[15 + k]; al[k] = Bi[18 + k]; }
            const double cm[3] = {tCom[3 * i], tCom[3 * i + 1], tCom[3 * i + 2]};
            const double In[3] = {tIn[3 * i], tIn[3 * i + 1], tIn[3 * i + 2]};
            const double m = tMass[i];
#pragma unroll
            for (int a = 0; a < 3; ++a) c[a] = R[3 * a] * cm[0] + R[3 * a + 1] * cm[1] + R[3 * a + 2] * cm[2];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    Iw[3 * a + b] = R[3 * a] * In[0] * R[3 * b] + R[3 * a + 1] * In[1] * R[3 * b + 1] + R[3 * a + 2] * In[2] * R[3 * b + 2];
            double t1[3], t2[3], t3[3], Iw_w[3], Iw_al[3], t4[3], fv[3], fw[3], pc[3], mom[3];
            cross3(al, c, t1);
            cross3(w, c, t2);
            cross3(w, t2, t3);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                Iw_w[a] = Iw[3 * a] * w[0] + Iw[3 * a + 1] * w[1] + Iw[3 * a + 2] * w[2];
                Iw_al[a] = Iw[3 * a] * al[0] + Iw[3 * a + 1] * al[1] + Iw[3 * a + 2] * al[2];
            }
            cross3(w, Iw_w, t4);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pc[a] = Bi[9 + a] + c[a];                         // COM position in the world
                fv[a] = m * (Bi[21 + a] + t1[a] + t3[a] + (a == 2 ? grav : 0.0));
                fw[a] = Iw_al[a] + t4[a];
            }
            cross3(pc, fv, mom);
            const double cc = pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2];
            double* C = Bi + 24;
            C[0] = m; C[1] = m * pc[0]; C[2] = m * pc[1]; C[3] = m * pc[2];
            C[4] = Iw[0] + m * (cc - pc[0] * pc[0]); C[5] = Iw[1] - m * pc[0] * pc[1]; C[6] = Iw[2] - m * pc[0] * pc[2];
            C[7] = Iw[4] + m * (cc - pc[1] * pc[1]); C[8] = Iw[5] - m * pc[1] * pc[2]; C[9] = Iw[8] + m * (cc - pc[2] * pc[2]);
            C[10] = fv[0]; C[11] = fv[1]; C[12] = fv[2];
            C[13] = fw[0] + mom[0]; C[14] = fw[1] + mom[1]; C[15] = fw[2] + mom[2];
        }
        __syncwarp();
        // composites: every body folded into its parent, children in descending index order (fixed order: no atomics);
        // lane k owns entry k of every body's block, so the sweep needs no synchronisation
        if (lane < 16) {
#pragma unroll 1
            for (int j = nb - 1; j >= 1; --j) Bd[RBD_BS * tParent[j] + 24 + lane] += Bd[RBD_BS * j + 24 + lane];
        }
        __syncwarp();
        // ---- 3. per column: twist about the world origin and the momentum of the subtree it moves; h
#pragma unroll 1
        for (int col = lane; col < nv; col += 32) {
            double om[3] = {0, 0, 0}, vo[3] = {0, 0, 0};
            const double p0[3] = {Bd[9], Bd[10], Bd[11]};
            int body = 0;
            if (col < 3) { vo[0] = col == 0 ? 1.0 : 0.0; vo[1] = col == 1 ? 1.0 : 0.0; vo[2] = col == 2 ? 1.0 : 0.0; }
            else if (col < 6) { om[0] = col == 3 ? 1.0 : 0.0; om[1] = col == 4 ? 1.0 : 0.0; om[2] = col == 5 ? 1.0 : 0.0; cross3(p0, om, vo); }
            else {
                body = col - 5;
                double pb[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) { om[k] = Bd[RBD_BS * body + 12 + k]; pb[k] = Bd[RBD_BS * body + 9 + k]; }
                cross3(pb, om, vo);
            }
            double C[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) C[k] = Bd[RBD_BS * body + 24 + k];
            const double mc[3] = {C[1], C[2], C[3]};
            double l[3], n[3], t[3];
            cross3(om, mc, t);
#pragma unroll
            for (int a = 0; a < 3; ++a) l[a] = C[0] * vo[a] + t[a];
            cross3(mc, vo, t);
            n[0] = t[0] + C[4] * om[0] + C[5] * om[1] + C[6] * om[2];
            n[1] = t[1] + C[5] * om[0] + C[7] * om[1] + C[8] * om[2];
            n[2] = t[2] + C[6] * om[0] + C[8] * om[1] + C[9] * om[2];
            double* Tc = T + RBD_TS * col;
#pragma unroll
            for (int a = 0; a < 3; ++a) { Tc[a] = om[a]; Tc[3 + a] = vo[a]; Tc[6 + a] = n[a]; Tc[9 + a] = l[a]; }
            rec[sh.off_h + col] = om[0] * C[13] + om[1] * C[14] + om[2] * C[15] + vo[0] * C[10] + vo[1] * C[11] + vo[2] * C[12];
        }
        __syncwarp();
        // M(a, b) = twist_b . momentum_a for related columns (a >= b): a lane per packed entry
        {
            const int nm = nv * (nv + 1) / 2;
#pragma unroll 1
            for (int e = lane; e < nm; e += 32) {
                int a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);        // row of packed entry e (fixed up below)
                while (a * (a + 1) / 2 > e) --a;
                while ((a + 1) * (a + 2) / 2 <= e) ++a;
                const int b = e - a * (a + 1) / 2;
                const unsigned long long anc_a = a >= 6 ? tAnc[a - 5] : 0ull;
                const bool related = b < 6 || a == b || ((anc_a >> (b < 6 ? 0 : b - 6)) & 1ull);
                const double* Tb = T + RBD_TS * b; const double* Ta = T + RBD_TS * a + 6;
                const double val = Tb[0] * Ta[0] + Tb[1] * Ta[1] + Tb[2] * Ta[2] + Tb[3] * Ta[3] + Tb[4] * Ta[4] + Tb[5] * Ta[5];
                rec[sh.off_M + e] = related ? val : 0.0;
            }
        }
        // ---- 4. task links: waist (body 0) then the contact links; a lane per Jacobian column
        const double* gains = S + sh.s_gains;
        const double lam_w = 100.0 * gains[0], lam2_w = 20.0 * gains[1], lam_p = 100.0 * gains[2], lam2_p = 20.0 * gains[3];
#pragma unroll 1
        for (int t = 0; t <= nc; ++t) {
            const int body = t == 0 ? 0 : tContact[t - 1];
            const unsigned long long an = tAnc[body];
            const double pt[3] = {Bd[RBD_BS * body + 9], Bd[RBD_BS * body + 10], Bd[RBD_BS * body + 11]}, p0[3] = {Bd[9], Bd[10], Bd[11]};
            double jvel[6] = {0, 0, 0, 0, 0, 0};
            double* Jout = rec + (t == 0 ? sh.off_jwaist : sh.off_jc + (t - 1) * 6 * nv);
#pragma unroll 1
            for (int col = lane; col < nv; col += 32) {
                double jv[3] = {0, 0, 0}, jw[3] = {0, 0, 0};
                if (col < 3) { jv[0] = col == 0 ? 1.0 : 0.0; jv[1] = col == 1 ? 1.0 : 0.0; jv[2] = col == 2 ? 1.0 : 0.0; }
                else if (col < 6) {
                    const double r[3] = {pt[0] - p0[0], pt[1] - p0[1], pt[2] - p0[2]};
                    jw[0] = col == 3 ? 1.0 : 0.0; jw[1] = col == 4 ? 1.0 : 0.0; jw[2] = col == 5 ? 1.0 : 0.0;
                    cross3(jw, r, jv);
                }
                else if ((an >> (col - 6)) & 1ull) {
                    const int b = col - 5;
                    double r[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { jw[k] = Bd[RBD_BS * b + 12 + k]; r[k] = pt[k] - Bd[RBD_BS * b + 9 + k]; }
                    cross3(jw, r, jv);
                }
                const double vc = col < 6 ? tw[col] : qd[col - 6];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    Jout[r * nv + col] = jv[r]; Jout[(3 + r) * nv + col] = jw[r];
                    jvel[r] = fma(jv[r], vc, jvel[r]); jvel[3 + r] = fma(jw[r], vc, jvel[3 + r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) jvel[r] += __shfl_xor_sync(0xffffffffu, jvel[r], o);
            if (lane < 6) {
                const int r = lane;
                double jvr = jvel[0];
#pragma unroll
                for (int k = 1; k < 6; ++k) jvr = r == k ? jvel[k] : jvr;
                rec[sh.off_jdqd + 6 * t + r] = r < 3 ? Bd[RBD_BS * body + 21 + r] : Bd[RBD_BS * body + 18 + r - 3];
                double e;
                if (t == 0) e = r < 3 ? S[sh.s_wpos + r] : S[sh.s_ori + r - 3];   // reference = initial - 0.1 z (ref:src/ForceAcc.cpp:181)
                else e = S[sh.s_foot + 6 * (t - 1) + r];
                rec[sh.off_rhs + 6 * t + r] = (t == 0 ? lam_w : lam_p) * e - (t == 0 ? lam2_w : lam2_p) * jvr;
            }
            if (t > 0) {
                if (sh.flags & QPPVM_FLAG_FRICTION_CONES) {
                    if (lane < 9) rec[sh.off_cone + 10 * (t - 1) + lane] = Bd[RBD_BS * body + lane];
                    else if (lane == 9) rec[sh.off_cone + 10 * (t - 1) + 9] = S[sh.s_mu + t - 1];
                }
                if (lane < 6) rec[sh.off_fbox + 6 * (t - 1) + lane] = lane < 2 ? -1000.0 : (lane == 2 ? 10.0 : 1000.0);   // ref:src/ForceAcc.cpp:75-76
            }
        }
        // postural right-hand side, torque limits, padding
        for (int j = lane; j < nv; j += 32) {
            const double e = j < 6 ? 0.0 : tQh[j - 6] - q[j - 6];
            const double vc = j < 6 ? tw[j] : qd[j - 6];
            rec[sh.off_rhs + 6 * (1 + nc) + j] = lam_p * e - lam2_p * vc;
        }
        if (sh.flags & QPPVM_FLAG_TORQUE_LIMITS)
            for (int a = lane; a < na; a += 32) {
                const double tm = tTm[a] * S[sh.s_tscale + a];
                rec[sh.off_taulim + a] = -tm;
                rec[sh.off_taulim + na + a] = tm;
            }
        for (int k = sh.off_fbox + 6 * nc + lane; k < sh.rec_doubles; k += 32) rec[k] = 0.0;
    }
}

// Command side: advance every state by one control period with the solved generalised
// acceleration -- the integration ref:src/ForceAcc.cpp:225-226 carries (commented out there; the plugin sends the
// model state as the position reference instead): q += dt*qd + dt^2/2*qdd, qd += dt*qdd; the floating base the same
// way in the world-aligned convention of the front end (position; rotation through the exponential of the world
// rotation vector dt*w + dt^2/2*alpha).  A state whose solve failed is left untouched (nothing is commanded,
// ref:src/ForceAcc.cpp:189-193).  One thread per state; 2 x state bytes + n_v doubles of traffic.
__global__ void __launch_bounds__(128)
integrate_states_kernel(RbdShape sh, double* __restrict__ states, const double* __restrict__ out,
                        const double* __restrict__ recs, int out_doubles, int trailer_off, double dt, long long batch)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch) return;
    double* st = states + idx * (size_t)sh.state_doubles;
    const double* x = out + idx * (size_t)out_doubles;
    if (reinterpret_cast<const int*>(x + trailer_off)[0] != 0) return;
    const double h2 = 0.5 * dt * dt;
    if (recs) {
        // The task references were captured once (ref:src/ForceAcc.cpp:158-164: resetReference at on_start, :181 waist
        // reference = initial - 0.1 z): the errors the states carry shrink as the links move.  Link velocity from the
        // record's right-hand side (rhs = lambda e - lambda2 J v), link acceleration J qdd + Jdot qdot.
        const double* rec = recs + idx * (size_t)sh.rec_doubles;
        const double* gains = st + sh.s_gains;
        const int nv = sh.n_v;
#pragma unroll 1
        for (int t = 0; t <= sh.n_c; ++t) {
            const double lam = 100.0 * gains[t == 0 ? 0 : 2], lam2 = 20.0 * gains[t == 0 ? 1 : 3];
            const double* J = rec + (t == 0 ? sh.off_jwaist : sh.off_jc + (t - 1) * 6 * nv);
#pragma unroll 1
            for (int r = 0; r < 6; ++r) {
                double* e = t == 0 ? (r < 3 ? st + sh.s_wpos + r : st + sh.s_ori + r - 3) : st + sh.s_foot + 6 * (t - 1) + r;
                double al = rec[sh.off_jdqd + 6 * t + r];
#pragma unroll 4
                for (int j = 0; j < nv; ++j) al = fma(J[r * nv + j], x[j], al);
                const double vl = (lam * *e - rec[sh.off_rhs + 6 * t + r]) / lam2;
                *e = *e - dt * vl - h2 * al;
            }
        }
    }
    double* tw = st + sh.s_tw;
    double th[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        st[sh.s_p0 + k] += dt * tw[k] + h2 * x[k];
        th[k] = dt * tw[3 + k] + h2 * x[3 + k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) tw[k] += dt * x[k];
    // R0 <- exp([th]x) R0  (Rodrigues; series for tiny angles)
    const double a2 = th[0] * th[0] + th[1] * th[1] + th[2] * th[2], ang = sqrt(a2);
    const double A = a2 > 1e-12 ? sin(ang) / ang : 1.0 - a2 / 6.0;
    const double Bc = a2 > 1e-12 ? (1.0 - cos(ang)) / a2 : 0.5 - a2 / 24.0;
    double R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = st[sh.s_R0 + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v[3] = {R[c], R[3 + c], R[6 + c]};
        double k1[3], k2[3];
        cross3(th, v, k1); cross3(th, k1, k2);
#pragma unroll
        for (int r = 0; r < 3; ++r) st[sh.s_R0 + 3 * r + c] = v[r] + A * k1[r] + Bc * k2[r];
    }
    const double* qdd = x + 6;
#pragma unroll 1
    for (int a = 0; a < sh.n_a; ++a) {
        const double v = st[sh.s_qd + a], acc = qdd[a];
        st[sh.s_q + a] += dt * v + h2 * acc;
        st[sh.s_qd + a] = v + dt * acc;
    }
}

}  // namespace qppvm
