// qp_kernel.cuh — batched hierarchical whole-body QP, one warp per problem, sm_100a.
//
// Replaces, per record, what one tick of the reference does between "model updated" and
// "torques written":
//   _autostack->update + QPOases_sot::solve (2 levels) + output recovery
//   ref:src/ForceAcc.cpp:184-219, ref:src/QPPVMPlugin.cpp:203-256  (SURVEY.md 8(a) a3-a16).
//
// B200 design (not the reference's: qpOASES is a sequential null-space homotopy method):
//   * one warp owns one QP; all state of the solve lives in that warp's shared-memory slab
//     (no H, C or KKT matrix ever touches HBM); the record is streamed once from HBM with
//     coalesced row reads; outputs are written once.
//   * whitening instead of normal equations: R from a Householder QR of the stacked task
//     matrix [sqrt(D+eps) ; A_dense] (never forms A^T A, so the eps-regularised directions
//     keep full relative accuracy), J = R^-1 explicit upper-triangular => every later
//     "solve" is a lane-parallel triangular mat-vec with no dependent chain across lanes.
//   * dual active set (Goldfarb-Idnani) in whitened coordinates u = R x, where the QP is a
//     least-distance problem; the active normals are kept as an orthonormal basis Q1 (CGS2)
//     plus a small triangular RN, so adding a constraint is two tall-skinny products, never
//     an n x n rotation sweep.  Equalities (dyn-feas, level-0 optimality rows) enter first.
//   * both priority levels, the qpOASES proximal regularisation re-solve, the KKT certificate
//     and tau = M qdd + h - J^T f run in the same kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qppvm_b200.h"

namespace qppvm {

constexpr int KMAX = 32;          // max simultaneously active constraints (eq + ineq)
constexpr int LDQ = KMAX + 1;     // row stride of Q1 (odd: conflict-free 64-bit column walks)
constexpr int LDR = KMAX + 1;     // column stride of RN

struct Params {
    double eps_reg;     // eps_regularisation * 2.221e-13
    int n_reg_steps;
    int max_iter;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// argmin over lanes of (v, idx); ties -> smaller idx.  Result uniform across the warp.
__device__ __forceinline__ void warp_argmin(double& v, int& idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// ------------------------------------------------------------------------------------------
// Problem policy: ForceAcc stack  x = [qddot ; f]   (ref:src/ForceAcc.cpp:58-137)
// ------------------------------------------------------------------------------------------
template <int NA_, int NC_, int FLAGS_>
struct ForceAcc {
    static constexpr int KIND = QPPVM_KIND_FORCEACC;
    static constexpr int NA = NA_, NC = NC_, FLAGS = FLAGS_;
    static constexpr bool CONES = (FLAGS & QPPVM_FLAG_FRICTION_CONES) != 0;
    static constexpr bool TLIM = (FLAGS & QPPVM_FLAG_TORQUE_LIMITS) != 0;
    static constexpr int NV = NA + 6, N = NV + 3 * NC;
    static constexpr int MD_MAX = 6 * NC > 6 ? 6 * NC : 6;   // dense task rows per level
    // reference row ids
    static constexpr int ROW_DYN = 0, ROW_BOX = 6, ROW_CONE = ROW_BOX + 6 * NC;
    static constexpr int ROW_TAU = ROW_CONE + (CONES ? 5 * NC : 0);
    static constexpr int ROW_OPT = ROW_TAU + (TLIM ? NA : 0);
    static constexpr int NROWS = ROW_OPT + QPPVM_M0;
    // inequality slots scanned each iteration
    static constexpr int NI_BOX = 3 * NC, NI_CONE = CONES ? 5 * NC : 0, NI_TAU = TLIM ? NA : 0;
    static constexpr int NI = NI_BOX + NI_CONE + NI_TAU;
    // record offsets (doubles)
    static constexpr int OFF_JW = 0, OFF_JC = OFF_JW + 6 * NV, OFF_M = OFF_JC + NC * 6 * NV;
    static constexpr int OFF_H = OFF_M + NV * (NV + 1) / 2, OFF_JDQD = OFF_H + NV;
    static constexpr int OFF_RHS = OFF_JDQD + 6 * (1 + NC), OFF_TAULIM = OFF_RHS + 6 * (1 + NC) + NV;
    static constexpr int OFF_CONE = OFF_TAULIM + (TLIM ? 2 * NA : 0);
    static constexpr int OFF_FBOX = OFF_CONE + (CONES ? 10 * NC : 0);
    static constexpr int REC_UNPADDED = OFF_FBOX + 6 * NC;
    static constexpr int REC = REC_UNPADDED + (REC_UNPADDED & 1);

    __device__ static __forceinline__ double M(const double* __restrict__ rec, int i, int j)
    {
        return i >= j ? __ldg(rec + OFF_M + i * (i + 1) / 2 + j) : __ldg(rec + OFF_M + j * (j + 1) / 2 + i);
    }
    __device__ static __forceinline__ int n_eq(int level) { return level == 0 ? 6 : 12; }
    __device__ static __forceinline__ int eq_row(int level, int e) { return e < 6 ? ROW_DYN + e : ROW_OPT + (e - 6); }
    __device__ static __forceinline__ bool regularised(int) { return true; }   // Cartesian/postural: HST_SEMIDEF

    // Dense task rows of a level into Ad (row-major, ld = N+1, last column = b); diagonal task
    // weights / targets into dg, db (postural rows are unit rows -> kept as a diagonal).
    // Level 0: waist Cartesian (ForceAcc.cpp:118-122).  Level 1: postural + contact Cartesian (:131).
    __device__ static int load_tasks(const double* __restrict__ rec, int level, double* Ad, double* dg, double* db, int lane)
    {
        constexpr int LDA = N + 1;
        const int md = level == 0 ? 6 : 6 * NC;
        for (int r = 0; r < md; ++r) {
            const double* Jr = level == 0 ? rec + OFF_JW + r * NV : rec + OFF_JC + r * NV;
            for (int j = lane; j < N; j += 32) Ad[r * LDA + j] = j < NV ? __ldg(Jr + j) : 0.0;
        }
        for (int r = lane; r < md; r += 32) {
            const int t = level == 0 ? r : 6 + r;            // task-row index into rhs / Jdqd
            Ad[r * LDA + N] = __ldg(rec + OFF_RHS + t) - __ldg(rec + OFF_JDQD + t);
        }
        for (int j = lane; j < N; j += 32) {
            const bool post = level == 1 && j < NV;
            dg[j] = post ? 1.0 : 0.0;
            db[j] = post ? __ldg(rec + OFF_RHS + 6 * (1 + NC) + j) : 0.0;
        }
        return md;
    }

    // Coefficients of constraint row `row` as a dense n-vector (smem av) + its two-sided bounds.
    // eopt: A0 x0* (level-1 optimality right-hand sides).
    __device__ static void build_row(const double* __restrict__ rec, int row, const double* eopt,
                                     double* av, double& lo, double& hi, int lane)
    {
        if (row < ROW_BOX) {                                   // DynamicFeasibility (base rows of M qdd + h - J^T w)
            const int r = row - ROW_DYN;
            for (int j = lane; j < N; j += 32) {
                double v;
                if (j < NV) v = M(rec, r, j);
                else { const int ci = (j - NV) / 3, k = (j - NV) % 3; v = -__ldg(rec + OFF_JC + (ci * 6 + k) * NV + r); }
                av[j] = v;
            }
            lo = hi = -__ldg(rec + OFF_H + r);
        } else if (row < ROW_CONE) {                           // wrench box (GenericConstraint)
            const int ci = (row - ROW_BOX) / 6, k = (row - ROW_BOX) % 6;
            for (int j = lane; j < N; j += 32) av[j] = (k < 3 && j == NV + 3 * ci + k) ? 1.0 : 0.0;
            if (k < 3) { lo = __ldg(rec + OFF_FBOX + 6 * ci + k); hi = __ldg(rec + OFF_FBOX + 6 * ci + 3 + k); }
            else { lo = -1.0; hi = 1.0; }
        } else if (CONES && row < ROW_TAU) {                   // friction pyramid on R^T f
            const int ci = (row - ROW_CONE) / 5, jr = (row - ROW_CONE) % 5;
            const double* R = rec + OFF_CONE + 10 * ci;
            const double mu = __ldg(R + 9) * 0.70710678118654752440;
            const double c0 = jr == 0 ? 1.0 : (jr == 1 ? -1.0 : 0.0);
            const double c1 = jr == 2 ? 1.0 : (jr == 3 ? -1.0 : 0.0);
            const double c2 = jr == 4 ? -1.0 : -mu;
            for (int j = lane; j < N; j += 32) {
                double v = 0.0;
                const int k = j - (NV + 3 * ci);
                if (k >= 0 && k < 3) v = c0 * __ldg(R + 3 * k) + c1 * __ldg(R + 3 * k + 1) + c2 * __ldg(R + 3 * k + 2);
                av[j] = v;
            }
            lo = -QPPVM_INFTY; hi = 0.0;
        } else if (TLIM && row < ROW_OPT) {                    // torque limits on M_a qdd + h_a - J_a^T f
            const int a = row - ROW_TAU;
            for (int j = lane; j < N; j += 32) {
                double v;
                if (j < NV) v = M(rec, 6 + a, j);
                else { const int ci = (j - NV) / 3, k = (j - NV) % 3; v = -__ldg(rec + OFF_JC + (ci * 6 + k) * NV + 6 + a); }
                av[j] = v;
            }
            const double ha = __ldg(rec + OFF_H + 6 + a);
            lo = __ldg(rec + OFF_TAULIM + a) - ha; hi = __ldg(rec + OFF_TAULIM + NA + a) - ha;
        } else {                                               // optimality rows of level 0: J_waist x = J_waist x0*
            const int r = row - ROW_OPT;
            for (int j = lane; j < N; j += 32) av[j] = j < NV ? __ldg(rec + OFF_JW + r * NV + j) : 0.0;
            lo = hi = eopt[r];
        }
    }

    // Inequality slot q -> (row id, value a.x, lo, hi).  One slot per lane.
    __device__ static void eval_slot(const double* __restrict__ rec, int q, const double* x,
                                     int& row, double& val, double& lo, double& hi)
    {
        if (q < NI_BOX) {
            const int ci = q / 3, k = q % 3;
            row = ROW_BOX + 6 * ci + k;
            val = x[NV + q];
            lo = __ldg(rec + OFF_FBOX + 6 * ci + k); hi = __ldg(rec + OFF_FBOX + 6 * ci + 3 + k);
        } else if (CONES && q < NI_BOX + NI_CONE) {
            const int qq = q - NI_BOX, ci = qq / 5, jr = qq % 5;
            row = ROW_CONE + qq;
            const double* R = rec + OFF_CONE + 10 * ci;
            const double mu = __ldg(R + 9) * 0.70710678118654752440;
            const double c0 = jr == 0 ? 1.0 : (jr == 1 ? -1.0 : 0.0);
            const double c1 = jr == 2 ? 1.0 : (jr == 3 ? -1.0 : 0.0);
            const double c2 = jr == 4 ? -1.0 : -mu;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                v += (c0 * __ldg(R + 3 * k) + c1 * __ldg(R + 3 * k + 1) + c2 * __ldg(R + 3 * k + 2)) * x[NV + 3 * ci + k];
            val = v; lo = -QPPVM_INFTY; hi = 0.0;
        } else {
            const int a = q - NI_BOX - NI_CONE;
            row = ROW_TAU + a;
            double v = 0.0;
            for (int j = 0; j < NV; ++j) v += M(rec, 6 + a, j) * x[j];
            for (int j = 0; j < 3 * NC; ++j) v -= __ldg(rec + OFF_JC + ((j / 3) * 6 + (j % 3)) * NV + 6 + a) * x[NV + j];
            const double ha = __ldg(rec + OFF_H + 6 + a);
            val = v; lo = __ldg(rec + OFF_TAULIM + a) - ha; hi = __ldg(rec + OFF_TAULIM + NA + a) - ha;
        }
    }

    // Level-0 task value A0 x0* (6 numbers) -> eopt.
    __device__ static void task0_value(const double* __restrict__ rec, const double* x, double* eopt, int lane)
    {
        for (int r = 0; r < QPPVM_M0; ++r) {
            double s = 0.0;
            for (int j = lane; j < NV; j += 32) s += __ldg(rec + OFF_JW + r * NV + j) * x[j];
            s = warp_sum(s);
            if (lane == 0) eopt[r] = s;
        }
    }

    // tau = (M qdd + h - sum J_c^T [f;0]) actuated rows  (ref:src/ForceAcc.cpp:206-219)
    __device__ static void recover(const double* __restrict__ rec, const double* x, double* tau_out, bool ok, int lane)
    {
        for (int a = lane; a < NA; a += 32) {
            double v = 0.0;
            if (ok) {
                v = __ldg(rec + OFF_H + 6 + a);
                for (int j = 0; j < NV; ++j) v += M(rec, 6 + a, j) * x[j];
                for (int j = 0; j < 3 * NC; ++j) v -= __ldg(rec + OFF_JC + ((j / 3) * 6 + (j % 3)) * NV + 6 + a) * x[NV + j];
            }
            tau_out[a] = v;     // failure: nothing is commanded (ForceAcc.cpp:189-193) -> zeros
        }
    }
};

// ------------------------------------------------------------------------------------------
// Per-warp shared-memory slab
// ------------------------------------------------------------------------------------------
template <class P>
struct Slab {
    static constexpr int N = P::N;
    static constexpr int LDJ = N | 1;                 // odd column stride
    static constexpr int LDA = N + 1;
    static constexpr int VEC = (N + 3) & ~3;
    static constexpr int SZ_J = N * LDJ;
    static constexpr int SZ_Q = (N * LDQ > P::MD_MAX * LDA) ? N * LDQ : P::MD_MAX * LDA;   // Q1, aliased by Ad
    static constexpr int SZ_R = KMAX * LDR;
    static constexpr int NVEC = 9;                    // u0 u x w w2 av dg db xp
    static constexpr int SZ_SMALL = 4 * KMAX + 8;     // d1 r lam | eopt
    static constexpr int DOUBLES = SZ_J + SZ_Q + SZ_R + NVEC * VEC + SZ_SMALL;
    static constexpr int INTS = 2 * KMAX + ((P::NROWS + 3) & ~3);   // act_row, act_sgn, cstate(bytes as ints/4)
    static constexpr int BYTES = DOUBLES * 8 + (2 * KMAX) * 4 + ((P::NROWS + 15) & ~15);
};

// ------------------------------------------------------------------------------------------
// The solver
// ------------------------------------------------------------------------------------------
template <class P>
struct Solver {
    using S = Slab<P>;
    static constexpr int N = P::N, LDJ = S::LDJ, LDA = S::LDA;

    double *Jm, *Q1, *Ad, *RN, *u0, *u, *x, *w, *w2, *av, *dg, *db, *xp, *d1, *rr, *lam, *eopt;
    int *act_row, *act_sgn;
    unsigned char* cstate;
    int lane, k, n_act_ineq, iters;
    const double* rec;

    __device__ void bind(unsigned char* slab, int lane_)
    {
        double* p = reinterpret_cast<double*>(slab);
        Jm = p; p += S::SZ_J;
        Q1 = p; Ad = p; p += S::SZ_Q;
        RN = p; p += S::SZ_R;
        u0 = p; p += S::VEC; u = p; p += S::VEC; x = p; p += S::VEC; w = p; p += S::VEC; w2 = p; p += S::VEC;
        av = p; p += S::VEC; dg = p; p += S::VEC; db = p; p += S::VEC; xp = p; p += S::VEC;
        d1 = p; p += KMAX; rr = p; p += KMAX; lam = p; p += KMAX; p += KMAX; eopt = p; p += 8;
        act_row = reinterpret_cast<int*>(p); act_sgn = act_row + KMAX;
        cstate = reinterpret_cast<unsigned char*>(act_sgn + KMAX);
        lane = lane_;
    }

    // ---- whitening: R^T R = D + eps I + Ad^T Ad via Householder QR of the stacked matrix, then J = R^-1.
    // u0 = R^-T (D db + Ad^T b + eps xp)  comes out as the transformed right-hand side.
    __device__ void factor(int md, double eps)
    {
        for (int i = lane; i < N * LDJ; i += 32) Jm[i] = 0.0;
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const double dd = dg[i] + eps;
            const double rt = sqrt(dd);
            Jm[i * LDJ + i] = rt;
            u0[i] = dd > 0.0 ? (dg[i] * db[i] + eps * xp[i]) / rt : 0.0;
        }
        __syncwarp();
        for (int kc = 0; kc < N; ++kc) {
            double sigma = 0.0;
            for (int r = 0; r < md; ++r) { const double a = Ad[r * LDA + kc]; sigma = fma(a, a, sigma); }
            if (sigma == 0.0) continue;                        // warp-uniform
            const double alpha = Jm[kc * LDJ + kc];
            const double nrm = sqrt(fma(alpha, alpha, sigma));
            const double v1 = -sigma / (alpha + nrm);          // alpha - nrm, cancellation-free (alpha >= 0)
            const double tau = 2.0 / fma(v1, v1, sigma);
            const double top_rhs = u0[kc];
            __syncwarp();
            for (int j = kc + 1 + lane; j <= N; j += 32) {     // trailing columns + rhs column (j == N)
                double s = (j == N) ? v1 * top_rhs : 0.0;
                for (int r = 0; r < md; ++r) s = fma(Ad[r * LDA + kc], Ad[r * LDA + j], s);
                s *= tau;
                if (j == N) u0[kc] = top_rhs - s * v1; else Jm[j * LDJ + kc] = -s * v1;
                for (int r = 0; r < md; ++r) Ad[r * LDA + j] = fma(-s, Ad[r * LDA + kc], Ad[r * LDA + j]);
            }
            if (lane == 0) Jm[kc * LDJ + kc] = nrm;
            __syncwarp();
        }
        // in-place inverse of the upper-triangular R (column-major): row i of J overwrites row i of R
        for (int i = N - 1; i >= 0; --i) {
            const double rinv = 1.0 / Jm[i * LDJ + i];
            double acc[2];
#pragma unroll
            for (int pss = 0; pss < 2; ++pss) {
                const int j = lane + 32 * pss;
                double a = (j == i) ? 1.0 : 0.0;
                if (j < N && j > i)
                    for (int l = i + 1; l <= j; ++l) a = fma(-Jm[l * LDJ + i], Jm[j * LDJ + l], a);
                acc[pss] = a * rinv;
            }
            __syncwarp();
#pragma unroll
            for (int pss = 0; pss < 2; ++pss) {
                const int j = lane + 32 * pss;
                if (j < N && j >= i) Jm[j * LDJ + i] = acc[pss];
            }
            __syncwarp();
        }
    }

    // w_out = sgn * J^T av   (lanes over j; column j of J is contiguous in i)
    __device__ void whiten(const double* a, double sgn, double* out)
    {
        for (int j = lane; j < N; j += 32) {
            double s0 = 0.0, s1 = 0.0;
            const double* col = Jm + j * LDJ;
            int i = 0;
            for (; i + 1 <= j; i += 2) { s0 = fma(col[i], a[i], s0); s1 = fma(col[i + 1], a[i + 1], s1); }
            if (i <= j) s0 = fma(col[i], a[i], s0);
            out[j] = sgn * (s0 + s1);
        }
        __syncwarp();
    }
    // x = J u  (lanes over i; row i of J strided by LDJ across j, consecutive across lanes)
    __device__ void unwhiten(const double* uu, double* xx)
    {
        for (int i = lane; i < N; i += 32) {
            double s0 = 0.0, s1 = 0.0;
            int j = i;
            for (; j + 1 < N; j += 2) { s0 = fma(Jm[j * LDJ + i], uu[j], s0); s1 = fma(Jm[(j + 1) * LDJ + i], uu[j + 1], s1); }
            if (j < N) s0 = fma(Jm[j * LDJ + i], uu[j], s0);
            xx[i] = s0 + s1;
        }
        __syncwarp();
    }
    __device__ double dot(const double* a, const double* b)
    {
        double s = 0.0;
        for (int i = lane; i < N; i += 32) s = fma(a[i], b[i], s);
        return warp_sum(s);
    }

    // d1 (+)= Q1^T v ; v -= Q1 d  (one Gram-Schmidt pass against the k active normals)
    __device__ void gs_pass(double* v, bool accumulate)
    {
        if (lane < k) {
            double s0 = 0.0, s1 = 0.0;
            int i = 0;
            for (; i + 1 < N; i += 2) { s0 = fma(Q1[i * LDQ + lane], v[i], s0); s1 = fma(Q1[(i + 1) * LDQ + lane], v[i + 1], s1); }
            if (i < N) s0 = fma(Q1[i * LDQ + lane], v[i], s0);
            rr[lane] = s0 + s1;                              // rr used as scratch for this pass' coefficients
            d1[lane] = accumulate ? d1[lane] + s0 + s1 : s0 + s1;
        }
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            double s = v[i];
            for (int c = 0; c < k; ++c) s = fma(-Q1[i * LDQ + c], rr[c], s);
            v[i] = s;
        }
        __syncwarp();
    }

    // r = RN^-1 d1 (back substitution; lane c holds component c)
    __device__ void solve_rn()
    {
        double dv = lane < k ? d1[lane] : 0.0;
        for (int c = k - 1; c >= 0; --c) {
            const double rc = __shfl_sync(0xffffffffu, dv, c) / RN[c * LDR + c];
            if (lane == c) dv = rc;
            else if (lane < c) dv = fma(-RN[c * LDR + lane], rc, dv);
        }
        if (lane < k) rr[lane] = dv;
        __syncwarp();
    }

    __device__ void drop(int l)
    {
        const int row = act_row[l];
        __syncwarp();
        if (lane == 0) cstate[row] = 0;
        // shift columns l+1.. of RN (and bookkeeping) one to the left
        for (int c = l; c < k - 1; ++c) {
            if (lane <= c + 1) RN[c * LDR + lane] = RN[(c + 1) * LDR + lane];
            __syncwarp();
        }
        {
            double lv = 0.0; int ar = 0, as = 0;
            if (lane >= l && lane < k - 1) { lv = lam[lane + 1]; ar = act_row[lane + 1]; as = act_sgn[lane + 1]; }
            __syncwarp();
            if (lane >= l && lane < k - 1) { lam[lane] = lv; act_row[lane] = ar; act_sgn[lane] = as; }
            __syncwarp();
        }
        // Givens: re-triangularise rows i, i+1 ; same rotation on columns i, i+1 of Q1
        for (int i = l; i < k - 1; ++i) {
            const double a = RN[i * LDR + i], b = RN[i * LDR + i + 1];
            const double h = hypot(a, b);
            const double c = h > 0.0 ? a / h : 1.0, s = h > 0.0 ? b / h : 0.0;
            __syncwarp();
            if (lane >= i && lane < k - 1) {
                const double ra = RN[lane * LDR + i], rb = RN[lane * LDR + i + 1];
                RN[lane * LDR + i] = c * ra + s * rb;
                RN[lane * LDR + i + 1] = -s * ra + c * rb;
            }
            for (int r = lane; r < N; r += 32) {
                const double qa = Q1[r * LDQ + i], qb = Q1[r * LDQ + i + 1];
                Q1[r * LDQ + i] = c * qa + s * qb;
                Q1[r * LDQ + i + 1] = -s * qa + c * qb;
            }
            __syncwarp();
        }
        --k; --n_act_ineq;
    }

    // Adds constraint `row` with sign sgn (normal sgn*a, already whitened into w), current slack sp <= 0.
    // Returns status; handles partial steps (drops) per Goldfarb-Idnani.
    __device__ int add_constraint(int row, int sgn, bool is_eq, double sp, double bound_abs, int max_iter)
    {
        double up = 0.0;
        const double ww = dot(w, w);
        for (;;) {
            if (iters >= max_iter) return QPPVM_STATUS_MAX_ITER;
            for (int i = lane; i < N; i += 32) w2[i] = w[i];
            __syncwarp();
            double nrm2 = ww;
            if (k > 0) {
                gs_pass(w2, false);
                gs_pass(w2, true);                            // CGS2: "twice is enough"
                nrm2 = dot(w2, w2);
            }
            const bool dependent = !(nrm2 > 1e-22 * ww) || k >= N;
            const bool full = k >= KMAX;
            double t1 = 1e300; int l = -1;
            if (n_act_ineq > 0) {
                solve_rn();
                double cand = 1e300; int ci = 0x7fffffff;
                if (lane < k && (act_sgn[lane] & 1) && rr[lane] > 0.0) { cand = lam[lane] / rr[lane]; ci = lane; }
                warp_argmin(cand, ci);
                if (ci != 0x7fffffff) { t1 = cand; l = ci; }
            }
            if (dependent && l < 0) {
                if (is_eq && -sp <= 1e-8 * fmax(1.0, bound_abs)) return QPPVM_STATUS_OK;   // redundant, consistent
                return QPPVM_STATUS_INFEASIBLE;
            }
            const double t2 = dependent ? 1e300 : -sp / nrm2;
            const double t = t1 < t2 ? t1 : t2;
            if (full && t == t2) return QPPVM_STATUS_NUMERIC;   // active-set capacity exhausted
            if (n_act_ineq > 0) { if (lane < k) lam[lane] -= t * rr[lane]; }
            up += t;
            if (!dependent) {
                for (int i = lane; i < N; i += 32) u[i] = fma(t, w2[i], u[i]);
                sp = fma(t, nrm2, sp);
            }
            __syncwarp();
            ++iters;
            if (t == t2) {                                    // full step: row becomes active
                const double nr = sqrt(nrm2), inv = 1.0 / nr;
                for (int i = lane; i < N; i += 32) Q1[i * LDQ + k] = w2[i] * inv;
                if (lane < k) RN[k * LDR + lane] = d1[lane];
                if (lane == 0) {
                    RN[k * LDR + k] = nr; lam[k] = up; act_row[k] = row; act_sgn[k] = is_eq ? 2 * sgn : sgn;
                    cstate[row] = 1;
                }
                __syncwarp();
                ++k; if (!is_eq) ++n_act_ineq;
                return QPPVM_STATUS_OK;
            }
            drop(l);                                          // partial step: blocking constraint leaves
        }
    }

    // Scan all inactive inequality slots at x; most violated -> (row, sgn, slack<0, |bound|).  Also max violation.
    __device__ bool scan(int& row, int& sgn, double& sp, double& babs)
    {
        double worst = 0.0; int widx = 0x7fffffff; int wsgn = 0; double wb = 0.0;
        for (int q = lane; q < ((P::NI + 31) & ~31); q += 32) {
            if (q < P::NI) {
                int r; double val, lo, hi;
                P::eval_slot(rec, q, x, r, val, lo, hi);
                if (!cstate[r]) {
                    const double tol = 1e-9 * fmax(1.0, fabs(val));
                    const double sl = val - lo, su = hi - val;
                    if (lo > -0.5 * QPPVM_INFTY && sl < -tol && sl < worst) { worst = sl; widx = r; wsgn = 1; wb = fabs(lo); }
                    if (hi < 0.5 * QPPVM_INFTY && su < -tol && su < worst) { worst = su; widx = r; wsgn = -1; wb = fabs(hi); }
                }
            }
        }
        double v = worst; int idx = widx;
        warp_argmin(v, idx);
        if (idx == 0x7fffffff) return false;
        // fetch sgn / bound from the winning lane (the lane whose (worst, widx) equals the winner)
        const unsigned m = __ballot_sync(0xffffffffu, widx == idx && worst == v);
        const int src = __ffs(m) - 1;
        row = idx; sp = v;
        sgn = __shfl_sync(0xffffffffu, wsgn, src);
        babs = __shfl_sync(0xffffffffu, wb, src);
        return true;
    }

    // One level: returns status.  On return x holds the level solution, lam/act_* the multipliers.
    __device__ int solve_level(int level, const Params& prm, float& kkt_out, double* ydiag)
    {
        const double eps = P::regularised(level) ? prm.eps_reg : 0.0;
        const int steps = eps > 0.0 ? prm.n_reg_steps : 0;
        for (int i = lane; i < N; i += 32) xp[i] = 0.0;
        for (int i = lane; i < P::NROWS; i += 32) cstate[i] = 0;
        __syncwarp();
        int md = P::load_tasks(rec, level, Ad, dg, db, lane);
        __syncwarp();
        factor(md, eps);                                      // Ad destroyed; Q1 may now alias it
        int status = QPPVM_STATUS_OK;
        k = 0; n_act_ineq = 0;
        for (int i = lane; i < N; i += 32) u[i] = u0[i];
        __syncwarp();
        // ---- equalities first (dyn-feas, then level-0 optimality rows), never dropped
        const int neq = P::n_eq(level);
        for (int e = 0; e < neq && status == QPPVM_STATUS_OK; ++e) {
            const int row = P::eq_row(level, e);
            double lo, hi;
            P::build_row(rec, row, eopt, av, lo, hi, lane);
            __syncwarp();
            whiten(av, 1.0, w);
            const double s = dot(w, u) - lo;
            const int sgn = s > 0.0 ? -1 : 1;
            if (sgn < 0) { for (int i = lane; i < N; i += 32) w[i] = -w[i]; __syncwarp(); }
            status = add_constraint(row, sgn, true, -fabs(s), fabs(lo), prm.max_iter);
        }
        // ---- inequalities + proximal regularisation steps
        for (int step = 0; status == QPPVM_STATUS_OK; ++step) {
            for (;;) {
                unwhiten(u, x);
                int row, sgn; double sp, babs;
                if (!scan(row, sgn, sp, babs)) break;
                double lo, hi;
                P::build_row(rec, row, eopt, av, lo, hi, lane);
                __syncwarp();
                whiten(av, (double)sgn, w);
                status = add_constraint(row, sgn, false, sp, babs, prm.max_iter);
                if (status != QPPVM_STATUS_OK) break;
            }
            if (status != QPPVM_STATUS_OK || step >= steps) break;
            // qpOASES solveRegularisedQP(): g <- g_orig - eps x_prev, i.e. u0 += delta, delta = eps J^T (x - xp_old);
            // same active set: u += (I - Q1 Q1^T) delta, lam -= RN^-1 Q1^T delta.
            for (int i = lane; i < N; i += 32) { av[i] = eps * (x[i] - xp[i]); xp[i] = x[i]; }
            __syncwarp();
            whiten(av, 1.0, w);
            for (int i = lane; i < N; i += 32) { u0[i] += w[i]; w2[i] = w[i]; }
            __syncwarp();
            if (k > 0) {
                gs_pass(w2, false);
                gs_pass(w2, true);
                solve_rn();
                if (lane < k) lam[lane] -= rr[lane];
                // (equality multipliers are re-derived at output time from u - u0); inequality ones must stay >= 0
                const bool neg = lane < k && (act_sgn[lane] & 1) && lam[lane] < 0.0;
                if (__any_sync(0xffffffffu, neg)) {
                    // rare: active set changes under the proximal shift -> cold restart of this step from u0
                    for (int i = lane; i < P::NROWS; i += 32) cstate[i] = 0;
                    k = 0; n_act_ineq = 0;
                    for (int i = lane; i < N; i += 32) u[i] = u0[i];
                    __syncwarp();
                    for (int e = 0; e < neq && status == QPPVM_STATUS_OK; ++e) {
                        const int row = P::eq_row(level, e);
                        double lo, hi;
                        P::build_row(rec, row, eopt, av, lo, hi, lane);
                        __syncwarp();
                        whiten(av, 1.0, w);
                        const double s = dot(w, u) - lo;
                        const int sgn = s > 0.0 ? -1 : 1;
                        if (sgn < 0) { for (int i = lane; i < N; i += 32) w[i] = -w[i]; __syncwarp(); }
                        status = add_constraint(row, sgn, true, -fabs(s), fabs(lo), prm.max_iter);
                    }
                    continue;
                }
            }
            for (int i = lane; i < N; i += 32) u[i] += w2[i];
            __syncwarp();
        }
        if (status == QPPVM_STATUS_OK) {                       // non-finite data must not reach the command
            bool bad = false;
            for (int i = lane; i < N; i += 32) bad |= !isfinite(x[i]);
            if (__any_sync(0xffffffffu, bad)) status = QPPVM_STATUS_NUMERIC;
        }
        if (status != QPPVM_STATUS_OK) { kkt_out = __int_as_float(0x7f800000); return status; }
        kkt_out = (float)kkt(level, md, eps, ydiag);
        return status;
    }

    // Signed multipliers y (qpOASES convention: > 0 active at lA, < 0 at uA) from u - u0 = sum lam_c w_c:
    // RN lam = Q1^T (u - u0).  Recomputed here so that equality multipliers are exact after all updates.
    __device__ void final_multipliers()
    {
        for (int i = lane; i < N; i += 32) w2[i] = u[i] - u0[i];
        __syncwarp();
        if (lane < k) {
            double s = 0.0;
            for (int i = 0; i < N; ++i) s = fma(Q1[i * LDQ + lane], w2[i], s);
            d1[lane] = s;
        }
        __syncwarp();
        solve_rn();                                           // rr = multipliers of the signed normals
    }

    // KKT certificate of the solved (regularised, proximal-shifted) level problem, SURVEY.md 8(c).
    __device__ double kkt(int level, int md, double eps, double* ydiag)
    {
        final_multipliers();
        // stationarity pieces need the original task rows again (Ad was overwritten by Q1)
        // grad = D (x - db) + Ad^T (Ad x - b) + eps (x - xp) - sum_c y_c a_c ; kept in w (grad) and w2 (H x)
        // step 1: constraint part, while Q1 is still alive we only need rr / act_*.
        for (int i = lane; i < N; i += 32) { w[i] = 0.0; }
        __syncwarp();
        double rprim = 0.0, rcomp = 0.0, cxmax = 0.0, ymax = 0.0;
        for (int c = 0; c < k; ++c) {
            const int row = act_row[c];
            const int sg = act_sgn[c];
            double lo, hi;
            P::build_row(rec, row, eopt, av, lo, hi, lane);
            __syncwarp();
            // rr[c] multiplies the signed whitened normal stored at insertion (sign = +-1, eq: +-2)
            const bool iseq = !(sg & 1);
            const double y = (sg > 0 ? 1.0 : -1.0) * rr[c];
            const double val = dot(av, x);
            for (int i = lane; i < N; i += 32) w[i] = fma(-y, av[i], w[i]);
            cxmax = fmax(cxmax, fabs(val)); ymax = fmax(ymax, fabs(y));
            if (iseq) rprim = fmax(rprim, fabs(val - lo));
            else {
                rprim = fmax(rprim, fmax(0.0, fmax(lo - val, val - hi)));
                rcomp = fmax(rcomp, y > 0.0 ? y * fabs(val - lo) : -y * fabs(hi - val));
                if (sg * y < 0.0) rcomp = fmax(rcomp, fabs(y));        // wrong-signed multiplier
                if (y == 0.0 && lane == 0) cstate[row] = 2;            // weakly active: not reported in the mask
            }
            if (ydiag && lane == 0) ydiag[row] = y;
            __syncwarp();
        }
        // inactive inequalities: primal violation only
        {
            double viol = 0.0, cm = 0.0;
            for (int q = lane; q < ((P::NI + 31) & ~31); q += 32)
                if (q < P::NI) {
                    int r; double val, lo, hi;
                    P::eval_slot(rec, q, x, r, val, lo, hi);
                    cm = fmax(cm, fabs(val));
                    if (!cstate[r]) viol = fmax(viol, fmax(lo - val, val - hi));
                }
            rprim = fmax(rprim, warp_max(viol)); cxmax = fmax(cxmax, warp_max(cm));
        }
        // step 2: task part (reload the dense task rows over the dead Q1 region)
        P::load_tasks(rec, level, Ad, dg, db, lane);
        __syncwarp();
        if (lane < md) {                                      // residual_r = Ad[r] x - b_r ; also keep (Ad x)_r
            double s = 0.0;
            for (int j = 0; j < N; ++j) s = fma(Ad[lane * LDA + j], x[j], s);
            d1[lane] = s;                                     // (A x)_r
            rr[lane] = Ad[lane * LDA + N];                    // b_r
        }
        __syncwarp();
        double rs = 0.0, gmax = 0.0, hxmax = 0.0, xmax = 0.0;
        for (int j = lane; j < N; j += 32) {
            double hx = (dg[j] + eps) * x[j], g = -dg[j] * db[j] - eps * xp[j];
            for (int r = 0; r < md; ++r) { hx = fma(Ad[r * LDA + j], d1[r], hx); g = fma(-Ad[r * LDA + j], rr[r], g); }
            const double st = hx + g + w[j];
            rs = fmax(rs, fabs(st)); gmax = fmax(gmax, fabs(g)); hxmax = fmax(hxmax, fabs(hx)); xmax = fmax(xmax, fabs(x[j]));
        }
        rs = warp_max(rs); gmax = warp_max(gmax); hxmax = warp_max(hxmax); xmax = warp_max(xmax);
        rs /= fmax(1.0, fmax(gmax, hxmax));
        rprim /= fmax(1.0, fmax(xmax, cxmax));
        rcomp /= fmax(1.0, ymax) * fmax(1.0, cxmax);
        return fmax(rs, fmax(rprim, rcomp));
    }
};

// ------------------------------------------------------------------------------------------
// Kernel: persistent warps pull problem indices from a global counter.
// ------------------------------------------------------------------------------------------
template <class P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
qp_solve_kernel(const double* __restrict__ recs, unsigned char* __restrict__ out, double* __restrict__ diag,
                long long batch, Params prm, unsigned long long* __restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using S = Slab<P>;
    constexpr int N = P::N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Solver<P> sv;
    sv.bind(smem + (size_t)warp * S::BYTES, lane);
    constexpr int OUT_BYTES = 8 * (N + P::NA) + 32;
    constexpr int DIAG = N + 2 * P::NROWS + QPPVM_M0;
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1ull);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if ((long long)idx >= batch) break;
        sv.rec = recs + idx * (size_t)P::REC;
        double* xo = reinterpret_cast<double*>(out + idx * (size_t)OUT_BYTES);
        double* dg = diag ? diag + idx * (size_t)DIAG : nullptr;
        if (dg) for (int i = lane; i < DIAG; i += 32) dg[i] = 0.0;
        sv.iters = 0;
        float kkt0 = __int_as_float(0x7f800000), kkt1 = kkt0;
        int it0 = 0, it1 = 0;
        int status = sv.solve_level(0, prm, kkt0, dg ? dg + N : nullptr);
        it0 = sv.iters;
        if (status == QPPVM_STATUS_OK) {
            P::task0_value(sv.rec, sv.x, sv.eopt, lane);
            __syncwarp();
            if (dg) {
                for (int i = lane; i < N; i += 32) dg[i] = sv.x[i];
                if (lane < QPPVM_M0) dg[N + 2 * P::NROWS + lane] = sv.eopt[lane];
            }
            sv.iters = 0;
            status = sv.solve_level(1, prm, kkt1, dg ? dg + N + P::NROWS : nullptr);
            it1 = sv.iters;
        }
        const bool ok = status == QPPVM_STATUS_OK;
        for (int i = lane; i < N; i += 32) xo[i] = ok ? sv.x[i] : 0.0;
        P::recover(sv.rec, sv.x, xo + N, ok, lane);
        // trailer: status, iters, 128-bit active mask of level 1, kkt[2]
        uint32_t mask = 0;
        if (ok && lane < 4)
            for (int b = 0; b < 32; ++b) {
                const int r = lane * 32 + b;
                if (r < P::NROWS && sv.cstate[r] == 1) mask |= 1u << b;
            }
        uint32_t* tr = reinterpret_cast<uint32_t*>(xo + N + P::NA);
        if (lane == 0) { tr[0] = (uint32_t)status; tr[1] = (uint32_t)((it0 & 0xffff) | (it1 << 16)); }
        if (lane < 4) tr[2 + lane] = mask;
        if (lane == 0) { tr[6] = __float_as_uint(kkt0); tr[7] = __float_as_uint(kkt1); }
        __syncwarp();
    }
}

}  // namespace qppvm
