// qp_kernel.cuh — batched hierarchical whole-body QP, one thread team per problem, sm_100a.
//
// Replaces, per record, what one tick of the reference does between "model updated" and
// "torques written":
//   _autostack->update + QPOases_sot::solve (2 levels) + output recovery
//   ref:src/ForceAcc.cpp:184-219, ref:src/QPPVMPlugin.cpp:203-256  (SURVEY.md 8(a) a3-a16).
//
// B200 design (not the reference's: qpOASES is a sequential null-space homotopy method):
//   * one team (TEAM = 64 threads = one CTA) owns one QP; the whole solve lives in the CTA's
//     shared-memory slab: no H, C or KKT matrix ever touches HBM.  Shapes whose inequality scan
//     re-reads M every iteration keep the re-read part of the record in shared memory (one TMA
//     bulk copy, cp.async.bulk + mbarrier); the others read the record through L1/L2 and spend
//     the shared memory on more resident CTAs.  Outputs are written once, coalesced.
//   * whitening instead of normal equations: R from a Householder QR of the stacked task
//     matrix [sqrt(D+eps) ; A_dense] (never forms A^T A, so the eps-regularised directions keep
//     full relative accuracy); only the leading NB x NB block that has dense task columns is
//     factorised, the cost-free force columns stay diagonal.  J = R^-1 is made explicit so
//     every later "solve" is a thread-parallel triangular mat-vec, no dependent chain across
//     threads.
//   * dual active set (Goldfarb-Idnani) in whitened coordinates u = R x, where the QP is a
//     least-distance problem; the active normals are kept as an orthonormal basis Q1 (CGS2)
//     plus a small triangular RN, so adding a constraint is two tall-skinny products, never an
//     n x n rotation sweep.  Equalities (dyn-feas, level-0 optimality rows) enter first.
//   * both priority levels, the qpOASES proximal regularisation re-solve, the KKT certificate
//     and tau = M qdd + h - J^T f run in the same kernel.
//   * ForceAcc shapes: the part of the above without data-dependent control flow -- the two
//     factorisations and the orthogonalisation of the equality rows -- runs in a separate, lane-batched
//     kernel (qp_factor_kernel: several (problem, level) pairs per CTA) and reaches the solve kernel
//     through a per-problem workspace in global memory that is laid out like the shared-memory slab and
//     fetched with bulk copies.  The Torque kind runs everything in qp_solve_kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qppvm_b200.h"

#ifndef QPPVM_SPLIT
#define QPPVM_SPLIT 1
#endif
#ifndef QPPVM_STAGE_BIG
#define QPPVM_STAGE_BIG 0        // 1: the 51-variable shapes also keep the record tail in shared memory (5 instead of 7 CTAs per SM)
#endif
#ifndef QPPVM_J_GLOBAL
#define QPPVM_J_GLOBAL 1
#endif
#ifndef QPPVM_SELECTIVE_GS
#define QPPVM_SELECTIVE_GS 1     // second Gram-Schmidt pass only when the first one cancelled more than half of the norm
#endif
#ifndef QPPVM_GS_RATIO
#define QPPVM_GS_RATIO 0.5       // ... i.e. when |w2|^2 < QPPVM_GS_RATIO |w|^2 after the first pass
#endif
#ifndef QPPVM_WS_COMPACT_Q
#define QPPVM_WS_COMPACT_Q 2     // the prepare workspace carries the n_eq filled columns of Q1 only (N x n_eq, contiguous), not the whole N x LDQ block;
                                 // 1: the solve kernel fetches them with plain loads, 2: with one bulk copy into the tail of the Q1 region + a scatter in shared memory
#endif
#ifndef QPPVM_PREP_INPLACE
#define QPPVM_PREP_INPLACE 1     // prepare kernel: the whitened equality normals overwrite the equality rows (4.9 KB less per pair: a 4th CTA per SM for the 51-variable shapes)
#endif
namespace qppvm {

// Active-set capacity KMAX (eq + ineq, <= 32 so one warp lane per active row), per problem shape.
__host__ __device__ constexpr int kmax_for(int n_eq, int n_ineq, int n)
{
    int k = n_eq + n_ineq;
    if (k > n) k = n;               // at most n linearly independent rows
    if (k > 32) k = 32;
    return k < 8 ? 8 : k;
}
constexpr int STATUS_IMPLIED = 100;   // internal: violated row is implied by the working set within tolerance

struct Params {
    double eps_reg;     // eps_regularisation * 2.221e-13
    int n_reg_steps;
    int max_iter;
    int rowwise;        // test switch (QPPVM_ROWWISE_EQUALITIES=1 at create): the prepare kernel flags every problem for
                        // the row-by-row equality path of the solve kernel, the one dependent rows fall back to
    // upstream-uncertain semantics made explicit (qppvm_desc), ForceAcc kind:
    double lam;         // lambda_solver: g = -lambda A^T W b
    double sw[3];       // square roots of the task weights: waist, postural, contact Cartesian (W = w I per task)
    int post_act_only;  // Postural without the six floating-base rows
};

// Latency mode (one QP per control tick, ref:src/QPPVMPlugin.cpp:308-329): the three kernels stay RESIDENT, one CTA
// each, and hand one problem along prepare -> solve -> certify through flags instead of being launched every tick.
// The host posts a tick by writing the record into pinned memory and bumping host[0]; the prepare server polls that
// word over PCIe, pulls the record into device memory, and the chain ends with the output record and host[1] = tick
// written back to pinned memory.  A null `host` means the ordinary batched launch.
struct Tick {
    volatile uint32_t* host;    // pinned host words: [0] tick posted, [1] tick completed, [2] quit request, [3] servers alive
    uint32_t* dev;              // device words: [0] prepare done, [1] solve done, [2] prepare server left, [3] solve server left
    const double* host_rec;     // the record, pinned host memory
    double* dev_rec;            // its copy in device memory (what the kernels read)
    double* host_out;           // the output record, pinned host memory
    uint32_t seq0;              // last tick completed before these servers were launched
    uint32_t idle_us;           // the prepare server leaves (and takes the chain down) after this long without a tick
};
__device__ __forceinline__ uint32_t ld_acquire_sys(const volatile uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(volatile uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stage timestamps of the resident chain (device words 8 ..: seven 64-bit stamps, see qppvm_tick_stamps)
__device__ __forceinline__ void tick_stamp(const Tick& tk, int slot)
{
    reinterpret_cast<unsigned long long*>(tk.dev + 8)[slot] = globaltimer_ns();
}
// Waits (one thread) until the device word differs from `last`; returns false when the upstream server has left.
__device__ __forceinline__ bool tick_wait_dev(const Tick& tk, int word, int down_word, uint32_t& last)
{
    for (;;) {
        const uint32_t v = ld_acquire_gpu(tk.dev + word);
        if (v != last) { last = v; return true; }
        if (ld_acquire_gpu(tk.dev + down_word) != 0u) {
            const uint32_t v2 = ld_acquire_gpu(tk.dev + word);  // a tick published just before the shutdown still counts
            if (v2 != last) { last = v2; return true; }
            return false;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Team primitives (TEAM threads = the whole CTA)
// ------------------------------------------------------------------------------------------
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
// argmin of (v, idx); ties -> smaller idx.  Result uniform across the warp.
__device__ __forceinline__ void warp_argmin(double& v, int& idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

extern __shared__ __align__(16) unsigned char g_smem[];   // the CTA's slab; all addresses are compile-time offsets

template <int TEAM>
struct Team {
    static constexpr int WARPS = TEAM / 32;
    __device__ static __forceinline__ void sync()
    {
        if (TEAM == 32) __syncwarp(); else __syncthreads();
    }
    // cross-warp exchange (TEAM > 32 only): red = 2 x WARPS doubles.
    // sum / argmin end WITHOUT a barrier after the read of `red`: the caller guarantees that every thread passes at
    // least one team barrier before the next reduction writes the exchange slots again (audited per call site: in the
    // active-set loop each reduction is followed by a phase that ends with a barrier of its own).  The one barrier
    // inside also publishes whatever the threads wrote to shared memory before the call.
    __device__ static __forceinline__ double sum(double v, double* red)
    {
        v = warp_sum(v);
        if (TEAM > 32) {
            const int tid = threadIdx.x;
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            v = red[0];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) v += red[w];
        }
        return v;
    }
    __device__ static __forceinline__ double max(double v, double* red)
    {
        v = warp_max(v);
        if (TEAM > 32) {
            const int tid = threadIdx.x;
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            v = red[0];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) v = fmax(v, red[w]);
            __syncthreads();
        }
        return v;
    }
    __device__ static __forceinline__ void argmin(double& v, int& idx, double* red)
    {
        warp_argmin(v, idx);
        if (TEAM > 32) {
            const int tid = threadIdx.x;
            if ((tid & 31) == 0) { red[tid >> 5] = v; red[WARPS + (tid >> 5)] = (double)idx; }
            __syncthreads();
            v = red[0]; idx = (int)red[WARPS];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) {
                const double ov = red[w]; const int oi = (int)red[WARPS + w];
                if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
            }
        }
    }
    __device__ static __forceinline__ bool any(bool p)
    {
        if (TEAM == 32) return __any_sync(0xffffffffu, p);
        return __syncthreads_or(p) != 0;
    }
};

// ---- TMA bulk copy global -> shared with mbarrier completion (UBLKCP in SASS) ---------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    // order the previous problem's generic-proxy reads of the slab before the async-proxy overwrite
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// one instruction: pull `bytes` (multiple of 16) at gsrc (16-byte aligned) into L2
__device__ __forceinline__ void l2_prefetch(const void* gsrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// the two halves of bulk_load, for several copies completing on one barrier phase
__device__ __forceinline__ void bulk_expect(uint64_t* bar, uint32_t bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------
// Problem policy: ForceAcc stack  x = [qddot ; f]   (ref:src/ForceAcc.cpp:58-137)
// `rec` points at the record staged in shared memory.
// ------------------------------------------------------------------------------------------
template <int NA_, int NC_, int FLAGS_>
struct ForceAcc {
    static constexpr int KIND = QPPVM_KIND_FORCEACC;
    static constexpr int NA = NA_, NC = NC_, FLAGS = FLAGS_;
    static constexpr bool CONES = (FLAGS & QPPVM_FLAG_FRICTION_CONES) != 0;
    static constexpr bool TLIM = (FLAGS & QPPVM_FLAG_TORQUE_LIMITS) != 0;
    // variables per contact: 3 force components, or the full wrench ("put 6 for full wrench", ref:src/ForceAcc.cpp:67)
    static constexpr int WD = (FLAGS & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3;
    static constexpr int NV = NA + 6, N = NV + WD * NC;
    // QPPVM_FLAG_COM_TASK: the centroidal force task (OpenSoT tasks::force::CoM, ref:src/ForceAcc.cpp:103) joins level 1.
    // Its rows act on the wrench variables, so those columns are dense in the level Hessian: the whitening covers all N
    // columns, every inequality row goes through the triangular products (no force-only shortcut), and the shape runs
    // the general in-kernel path (factorisation, row-by-row equalities, KKT check) the Torque kind uses.
    static constexpr bool COM = (FLAGS & QPPVM_FLAG_COM_TASK) != 0;
    static constexpr int NB = COM ? N : NV;                  // columns with dense task entries (both levels)
    static constexpr int MD1_ = 6 * NC + (COM ? 6 : 0);
    static constexpr int MD_MAX = MD1_ > 6 ? MD1_ : 6;       // dense task rows per level
    // reference row ids
    static constexpr int ROW_DYN = 0, ROW_BOX = 6, ROW_CONE = ROW_BOX + 6 * NC;
    static constexpr int ROW_TAU = ROW_CONE + (CONES ? 5 * NC : 0);
    static constexpr int ROW_OPT = ROW_TAU + (TLIM ? NA : 0);
    static constexpr int NROWS = ROW_OPT + QPPVM_M0;
    // inequality slots scanned each iteration
    static constexpr int NI_BOX = WD * NC, NI_CONE = CONES ? 5 * NC : 0, NI_TAU = TLIM ? NA : 0;
    static constexpr int NI = NI_BOX + NI_CONE + NI_TAU;
    // Slots [0, NI_FORCE) are rows over the force variables of ONE contact (box, friction pyramid).  Without a force
    // task the force columns carry no task entries, so they stay diagonal in the whitening (x_f = jd u_f): such a row is
    // evaluated from u without the triangular product, and its whitened normal has three entries ("cheap" slots).  The
    // torque-limit rows are dense.
    static constexpr int NI_FORCE = NI_BOX + NI_CONE;
    static constexpr int NI_CHEAP = COM ? 0 : NI_FORCE;
    // The loop over the working-set changes ends with a scan of the dense slots at the final x: their values
    // (M_a qdd - J_a^T f) are kept and reused by the KKT check and the torque recovery.
    static constexpr bool TAUVAL = TLIM && !COM;
    static constexpr int NTV = TLIM ? NA + (NA & 1) : 0;
    // record offsets (doubles)
    __host__ __device__ static constexpr int OFF_M_() { return 6 * (NA_ + 6) + NC_ * 6 * (NA_ + 6); }
    __host__ __device__ static constexpr int REC_()
    {
        const int nv = NA_ + 6;
        const int u = OFF_M_() + nv * (nv + 1) / 2 + nv + 6 * (1 + NC_) + 6 * (1 + NC_) + nv
                      + ((FLAGS_ & QPPVM_FLAG_TORQUE_LIMITS) ? 2 * NA_ : 0) + ((FLAGS_ & QPPVM_FLAG_FRICTION_CONES) ? 10 * NC_ : 0) + 2 * ((FLAGS_ & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3) * NC_
                      + ((FLAGS_ & QPPVM_FLAG_COM_TASK) ? 6 * ((FLAGS_ & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3) * NC_ + 6 : 0);
        return u + (u & 1);
    }
    static constexpr int OFF_JW = 0, OFF_JC = OFF_JW + 6 * NV, OFF_M = OFF_JC + NC * 6 * NV;
    static constexpr int OFF_H = OFF_M + NV * (NV + 1) / 2, OFF_JDQD = OFF_H + NV;
    static constexpr int OFF_RHS = OFF_JDQD + 6 * (1 + NC), OFF_TAULIM = OFF_RHS + 6 * (1 + NC) + NV;
    static constexpr int OFF_CONE = OFF_TAULIM + (TLIM ? 2 * NA : 0);
    static constexpr int OFF_FBOX = OFF_CONE + (CONES ? 10 * NC : 0);
    static constexpr int OFF_COM = OFF_FBOX + 2 * WD * NC;     // A_com 6 x (WD NC) | b_com 6
    static constexpr int REC_UNPADDED = OFF_COM + (COM ? 6 * WD * NC + 6 : 0);
    static constexpr int REC = REC_UNPADDED + (REC_UNPADDED & 1);
    static_assert(REC == REC_() && OFF_M == OFF_M_(), "layout helpers agree");

    __device__ static __forceinline__ double M(const double* rec, int i, int j)
    {
        return i >= j ? rec[OFF_M - SB + i * (i + 1) / 2 + j] : rec[OFF_M - SB + j * (j + 1) / 2 + i];
    }
    static constexpr int MD0 = 6, MD1 = MD1_;                  // dense task rows of level 0 / 1
    static constexpr int EXTRA = 0;                            // policy scratch in the slab (doubles)
    // The two factorisations depend only on the record: for the ForceAcc shapes they run in their own kernel
    // (qp_factor_kernel) and reach the solve through a workspace in global memory (L2-resident), which takes the
    // 25 KB of unrolled factorisation code out of the instruction-fetch-bound solve kernel (profiles/README.md).
    static constexpr bool SPLIT_FACTOR = QPPVM_SPLIT && !COM;
    // Staging (see Slab::STAGE): shapes whose inequality scan re-reads M every iteration (torque-limit rows) keep
    // the TAIL of the record [M | h | Jdqd | rhs | tau limits | cones | boxes] (one TMA bulk copy) plus the linear
    // contact-Jacobian rows in shared memory; the task Jacobians (read once per level) stay in global memory.
    // Policy functions get `rec` = staged tail (or the global record when nothing is staged) and `g` = global record.
    static constexpr bool STAGE_RECORD = TLIM && QPPVM_STAGE_BIG >= (NA_ + 6 + ((FLAGS_ & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3) * NC_ > 48);
    // J = R^-1 is only needed by the triangular products (full x for the dense slots / the end of a level, whitening
    // of a dense row): a few times per level since the force-only rows bypass them.  The 51-variable shape reads it
    // from the prepare workspace in global memory (L2) and spends the 6 KB on two more resident CTAs per SM.
    static constexpr bool J_GLOBAL = SPLIT_FACTOR && TLIM && !STAGE_RECORD && QPPVM_J_GLOBAL;
    static constexpr bool EXT_IS_GLOBAL = true;                // the `ext` argument carries the global record pointer
    static constexpr int STAGE_FROM = OFF_M_();                // first staged record offset (even => 16-byte aligned)
    static constexpr int SB = STAGE_RECORD ? STAGE_FROM : 0;   // staged offset = record offset - SB
    static constexpr int S_JCL = REC_() - STAGE_FROM;          // linear rows of J_c: (ci * 3 + k) * NV + col
    static constexpr int STAGED = S_JCL + WD * NC * NV + ((S_JCL + WD * NC * NV) & 1);
    __device__ static __forceinline__ double jcl(const double* rec, const double* g, int ci, int k, int col)
    {
        return STAGE_RECORD ? rec[S_JCL + (ci * WD + k) * NV + col] : g[OFF_JC + (ci * 6 + k) * NV + col];
    }
    static constexpr int KMAX_RAW = kmax_for(12, WD * NC + (CONES ? 5 * NC : 0) + (TLIM ? NA : 0), N);
    // 31 instead of 32 rows of capacity is what lets a fifth CTA of the 51-variable shape fit on an SM
    // (dense-force shapes: the helper lanes of the triangular products exchange through KP >= (NB - 1) / 2 slots)
    static constexpr int KMAX = (KMAX_RAW == 32 && N > 48) ? 31 : ((COM && KMAX_RAW < (N + 1) / 2) ? (N + 1) / 2 : KMAX_RAW);
    template <int TEAM> __device__ static __forceinline__ bool prepare(const double*, double*, int) { return true; }
    __host__ __device__ static constexpr int n_eq(int level) { return level == 0 ? 6 : 12; }
    __device__ static __forceinline__ int eq_row(int level, int e) { return e < 6 ? ROW_DYN + e : ROW_OPT + (e - 6); }
    __device__ static __forceinline__ bool regularised(int) { return true; }   // Cartesian/postural: HST_SEMIDEF
    // Coefficient jc of equality row e as (record offset, sign), -1: zero.  Same rows as build_row (dyn-feas base
    // rows, then the level-0 task rows); used by the prepare kernel of the unstaged shapes.
    __device__ static __forceinline__ int eq_coef(int, int e, int jc, double& sg)
    {
        if (e < 6) {
            if (jc < NV) { sg = 1.0; return OFF_M + (e >= jc ? e * (e + 1) / 2 + jc : jc * (jc + 1) / 2 + e); }
            const int ci = (jc - NV) / WD, k = (jc - NV) % WD;
            sg = -1.0;
            return OFF_JC + (ci * 6 + k) * NV + e;
        }
        sg = 1.0;
        return jc < NV ? OFF_JW + (e - 6) * NV + jc : -1;
    }
    static constexpr bool HAS_EQ_COEF = true;
    // bound of equality row e as the solve kernel sees it (rec: staged tail or global record; eopt: level-0 task value)
    __device__ static __forceinline__ double eq_bound(const double* rec, int e, const double* eopt)
    {
        return e < 6 ? -rec[OFF_H - SB + e] : eopt[e - 6];
    }
    // right-hand side of equality row e when it depends on the record alone (dyn-feas rows)
    __device__ static __forceinline__ double eq_rhs(const double* g, int, int e) { return e < 6 ? -g[OFF_H + e] : 0.0; }

    // Dense task rows of a level into Ad (row-major, ld = NB+1, last column = b); diagonal task
    // weights / targets into dg, db (postural rows are unit rows -> kept as a diagonal).
    // Level 0: waist Cartesian (ForceAcc.cpp:118-122).  Level 1: postural + contact Cartesian (:131).
    // The cost of a level is sum_k w_k/2 ||A_k x - lambda b_k||^2 (H = A^T W A, g = -lambda A^T W b: SURVEY App. A.2):
    // rows are stored scaled by sqrt(w_k), right-hand sides by lambda sqrt(w_k).
    template <int TEAM>
    __device__ static int load_tasks(const double* rec, const double* g, int level, double* Ad, double* dg, double* db, int tid,
                                     const Params& prm)
    {
        constexpr int LDA = NB + 1;
        const int md = level == 0 ? 6 : 6 * NC;
        const double sw = level == 0 ? prm.sw[0] : prm.sw[2];
        const double* Jr = level == 0 ? g + OFF_JW : g + OFF_JC;
        for (int e = tid; e < md * NV; e += TEAM) { const int r = e / NV, j = e - r * NV; Ad[r * LDA + j] = sw * Jr[e]; }
        for (int r = tid; r < md; r += TEAM) {
            const int t = level == 0 ? r : 6 + r;            // task-row index into rhs / Jdqd
            Ad[r * LDA + NB] = sw * prm.lam * (rec[OFF_RHS - SB + t] - rec[OFF_JDQD - SB + t]);
        }
        if constexpr (COM) {
            // the Cartesian rows have no entries on the wrench columns; at level 1 the six CoM rows follow them
            // (zero on the acceleration columns, A_com on the wrench columns, right-hand side lambda b_com)
            constexpr int NF = WD * NC;
            for (int e = tid; e < md * NF; e += TEAM) { const int r = e / NF, j = e - r * NF; Ad[r * LDA + NV + j] = 0.0; }
            if (level == 1) {
                for (int e = tid; e < 6 * NV; e += TEAM) { const int r = e / NV, j = e - r * NV; Ad[(md + r) * LDA + j] = 0.0; }
                for (int e = tid; e < 6 * NF; e += TEAM) { const int r = e / NF, j = e - r * NF; Ad[(md + r) * LDA + NV + j] = rec[OFF_COM - SB + e]; }
                if (tid < 6) Ad[(md + tid) * LDA + NB] = prm.lam * rec[OFF_COM - SB + 6 * NF + tid];
            }
        }
        for (int j = tid; j < N; j += TEAM) {
            const bool post = level == 1 && j < NV && !(prm.post_act_only && j < 6);
            dg[j] = post ? prm.sw[1] * prm.sw[1] : 0.0;
            db[j] = post ? prm.lam * rec[OFF_RHS - SB + 6 * (1 + NC) + j] : 0.0;
        }
        return (COM && level == 1) ? md + 6 : md;
    }

    // Coefficients of constraint row `row` as a dense n-vector (smem av) + its two-sided bounds.
    // eopt: A0 x0* (level-1 optimality right-hand sides).
    template <int TEAM>
    __device__ static void build_row(const double* rec, const double* g, int row, const double* eopt,
                                     double* av, double& lo, double& hi, int tid)
    {
        if (row < ROW_BOX) {                                   // DynamicFeasibility (base rows of M qdd + h - J^T w)
            const int r = row - ROW_DYN;
            for (int j = tid; j < N; j += TEAM) {
                double v;
                if (j < NV) v = M(rec, r, j);
                else { const int ci = (j - NV) / WD, k = (j - NV) % WD; v = -jcl(rec, g, ci, k, r); }
                av[j] = v;
            }
            lo = hi = -rec[OFF_H - SB + r];
        } else if (row < ROW_CONE) {                           // wrench box (GenericConstraint)
            const int ci = (row - ROW_BOX) / 6, k = (row - ROW_BOX) % 6;
            for (int j = tid; j < N; j += TEAM) av[j] = (k < WD && j == NV + WD * ci + k) ? 1.0 : 0.0;
            if (k < WD) { lo = rec[OFF_FBOX - SB + 2 * WD * ci + k]; hi = rec[OFF_FBOX - SB + 2 * WD * ci + WD + k]; }
            else { lo = -1.0; hi = 1.0; }
        } else if (CONES && row < ROW_TAU) {                   // friction pyramid on R^T f
            const int ci = (row - ROW_CONE) / 5, jr = (row - ROW_CONE) % 5;
            const double* R = rec + OFF_CONE - SB + 10 * ci;
            const double mu = R[9] * 0.70710678118654752440;
            const double c0 = jr == 0 ? 1.0 : (jr == 1 ? -1.0 : 0.0);
            const double c1 = jr == 2 ? 1.0 : (jr == 3 ? -1.0 : 0.0);
            const double c2 = jr == 4 ? -1.0 : -mu;
            for (int j = tid; j < N; j += TEAM) {
                double v = 0.0;
                const int k = j - (NV + WD * ci);
                if (k >= 0 && k < 3) v = c0 * R[3 * k] + c1 * R[3 * k + 1] + c2 * R[3 * k + 2];
                av[j] = v;
            }
            lo = -QPPVM_INFTY; hi = 0.0;
        } else if (TLIM && row < ROW_OPT) {                    // torque limits on M_a qdd + h_a - J_a^T f
            const int a = row - ROW_TAU;
            for (int j = tid; j < N; j += TEAM) {
                double v;
                if (j < NV) v = M(rec, 6 + a, j);
                else { const int ci = (j - NV) / WD, k = (j - NV) % WD; v = -jcl(rec, g, ci, k, 6 + a); }
                av[j] = v;
            }
            const double ha = rec[OFF_H - SB + 6 + a];
            lo = rec[OFF_TAULIM - SB + a] - ha; hi = rec[OFF_TAULIM - SB + NA + a] - ha;
        } else {                                               // optimality rows of level 0: J_waist x = J_waist x0*
            const int r = row - ROW_OPT;
            for (int j = tid; j < N; j += TEAM) av[j] = j < NV ? g[OFF_JW + r * NV + j] : 0.0;
            lo = hi = eopt[r];
        }
    }

    // Slot table (shared memory, filled once per problem): per force-only slot its three coefficients on the force
    // variables of its contact and the two bounds.
    static constexpr int CT = 5;
    static constexpr int NCT = NI_FORCE * CT + ((NI_FORCE * CT) & 1);
    __device__ static __forceinline__ int slot_row(int q) { return q < NI_BOX ? ROW_BOX + 6 * (q / WD) + q % WD : ROW_CONE + (q - NI_BOX); }
    // first of the three consecutive variables the slot's coefficients refer to (force or torque part of a wrench)
    __device__ static __forceinline__ int slot_col(int q) { return q < NI_BOX ? NV + WD * (q / WD) + 3 * ((q % WD) / 3) : NV + WD * ((q - NI_BOX) / 5); }
    template <int TEAM>
    __device__ static __forceinline__ void fill_slot_table(const double* rec, double* ct, int tid)
    {
        for (int q = tid; q < NI_FORCE; q += TEAM) {
            double a0, a1, a2, lo, hi;
            if (q < NI_BOX) {                                  // wrench box (GenericConstraint), force rows
                const int ci = q / WD, k = q % WD;
                a0 = k % 3 == 0 ? 1.0 : 0.0; a1 = k % 3 == 1 ? 1.0 : 0.0; a2 = k % 3 == 2 ? 1.0 : 0.0;
                lo = rec[OFF_FBOX - SB + 2 * WD * ci + k]; hi = rec[OFF_FBOX - SB + 2 * WD * ci + WD + k];
            } else {                                           // friction pyramid on R^T f
                const int qq = q - NI_BOX, ci = qq / 5, jr = qq % 5;
                const double* R = rec + OFF_CONE - SB + 10 * ci;
                const double mu = R[9] * 0.70710678118654752440;
                const double c0 = jr == 0 ? 1.0 : (jr == 1 ? -1.0 : 0.0);
                const double c1 = jr == 2 ? 1.0 : (jr == 3 ? -1.0 : 0.0);
                const double c2 = jr == 4 ? -1.0 : -mu;
                a0 = c0 * R[0] + c1 * R[1] + c2 * R[2];
                a1 = c0 * R[3] + c1 * R[4] + c2 * R[5];
                a2 = c0 * R[6] + c1 * R[7] + c2 * R[8];
                lo = -QPPVM_INFTY; hi = 0.0;
            }
            double* t = ct + q * CT;
            t[0] = a0; t[1] = a1; t[2] = a2; t[3] = lo; t[4] = hi;
        }
    }
    // Force-only row (box rows with k < 3, pyramid rows): first column and the three coefficients.
    __device__ static __forceinline__ bool sparse_row(const double* ct, int row, int& j0, double& a0, double& a1, double& a2)
    {
        if (COM || row < ROW_BOX || row >= ROW_TAU) return false;     // (without cones ROW_TAU == ROW_CONE)
        const int q = row < ROW_CONE ? WD * ((row - ROW_BOX) / 6) + (row - ROW_BOX) % 6 : NI_BOX + (row - ROW_CONE);
        j0 = slot_col(q);
        a0 = ct[q * CT]; a1 = ct[q * CT + 1]; a2 = ct[q * CT + 2];
        return true;
    }

    // Inequality slot q -> (row id, value a.x, lo, hi).  One slot per thread.
    __device__ static void eval_slot(const double* rec, const double* g, const double* ct, int q, const double* x,
                                     int& row, double& val, double& lo, double& hi)
    {
        if (q < NI_FORCE) {
            const double* t = ct + q * CT;
            const double* xf = x + slot_col(q);
            row = slot_row(q);
            val = fma(t[0], xf[0], fma(t[1], xf[1], t[2] * xf[2]));
            lo = t[3]; hi = t[4];
        } else {
            const int a = q - NI_FORCE;
            row = ROW_TAU + a;
            const int i = 6 + a;
            const double* Mi = rec + OFF_M - SB + i * (i + 1) / 2;
            // four accumulators: the loads are independent, the record may sit in global memory (unstaged shapes)
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
            int j = 0;
#pragma unroll 1
            for (; j + 3 <= i; j += 4) {
                v0 = fma(Mi[j], x[j], v0); v1 = fma(Mi[j + 1], x[j + 1], v1);
                v2 = fma(Mi[j + 2], x[j + 2], v2); v3 = fma(Mi[j + 3], x[j + 3], v3);
            }
#pragma unroll 1
            for (; j <= i; ++j) v0 = fma(Mi[j], x[j], v0);
#pragma unroll 1
            for (j = i + 1; j + 1 < NV; j += 2) {
                v1 = fma(rec[OFF_M - SB + j * (j + 1) / 2 + i], x[j], v1);
                v3 = fma(rec[OFF_M - SB + (j + 1) * (j + 2) / 2 + i], x[j + 1], v3);
            }
            if (j < NV) v1 = fma(rec[OFF_M - SB + j * (j + 1) / 2 + i], x[j], v1);
#pragma unroll 1
            for (int ci = 0; ci < NC; ++ci)
#pragma unroll
                for (int k = 0; k < WD; ++k) v2 = fma(-jcl(rec, g, ci, k, i), x[NV + WD * ci + k], v2);
            const double ha = rec[OFF_H - SB + i];
            val = (v0 + v1) + (v2 + v3); lo = rec[OFF_TAULIM - SB + a] - ha; hi = rec[OFF_TAULIM - SB + NA + a] - ha;
        }
    }

    // ---- accessors of the certificate kernel (qp_certify_kernel): everything from the record in global memory ----
    // a_row . x and the bounds of constraint row `row` at `level`; false: the row does not exist at that level or is
    // one of the trivial rows 0 . x in [-1, 1] of the 6-row wrench boxes.
    __device__ static bool row_value(const double* g, int level, int row, const double* x, const double* eopt,
                                     double& val, double& lo, double& hi)
    {
        if (row < ROW_BOX || (TLIM && row >= ROW_TAU && row < ROW_OPT)) {      // row i of M qdd - J_c^T f
            const int i = row < ROW_BOX ? row : 6 + (row - ROW_TAU);
            const double* Mi = g + OFF_M + i * (i + 1) / 2;
            double v0 = 0.0, v1 = 0.0;
#pragma unroll 1
            for (int j = 0; j <= i; ++j) v0 = fma(Mi[j], x[j], v0);
#pragma unroll 1
            for (int j = i + 1; j < NV; ++j) v1 = fma(g[OFF_M + j * (j + 1) / 2 + i], x[j], v1);
#pragma unroll 1
            for (int f = 0; f < WD * NC; ++f) v0 = fma(-g[OFF_JC + ((f / WD) * 6 + f % WD) * NV + i], x[NV + f], v0);
            val = v0 + v1;
            const double hv = g[OFF_H + i];
            if (row < ROW_BOX) lo = hi = -hv;
            else { lo = g[OFF_TAULIM + row - ROW_TAU] - hv; hi = g[OFF_TAULIM + NA + row - ROW_TAU] - hv; }
            return true;
        }
        if (row < ROW_CONE) {
            const int ci = (row - ROW_BOX) / 6, k = (row - ROW_BOX) % 6;
            if (k >= WD) return false;
            val = x[NV + WD * ci + k]; lo = g[OFF_FBOX + 2 * WD * ci + k]; hi = g[OFF_FBOX + 2 * WD * ci + WD + k];
            return true;
        }
        if (CONES && row < ROW_TAU) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) v = fma(row_coef(g, row, NV + WD * ((row - ROW_CONE) / 5) + k), x[NV + WD * ((row - ROW_CONE) / 5) + k], v);
            val = v; lo = -QPPVM_INFTY; hi = 0.0;
            return true;
        }
        if (level == 0) return false;
        const double* Jr = g + OFF_JW + (row - ROW_OPT) * NV;
        double v0 = 0.0;
#pragma unroll 1
        for (int j = 0; j < NV; ++j) v0 = fma(Jr[j], x[j], v0);
        val = v0; lo = hi = eopt[row - ROW_OPT];
        return true;
    }
    // coefficient of variable j in constraint row `row`
    __device__ static double row_coef(const double* g, int row, int j)
    {
        if (row < ROW_BOX || (TLIM && row >= ROW_TAU && row < ROW_OPT)) {
            const int i = row < ROW_BOX ? row : 6 + (row - ROW_TAU);
            if (j < NV) return i >= j ? g[OFF_M + i * (i + 1) / 2 + j] : g[OFF_M + j * (j + 1) / 2 + i];
            const int f = j - NV;
            return -g[OFF_JC + ((f / WD) * 6 + f % WD) * NV + i];
        }
        if (row < ROW_CONE) {
            const int ci = (row - ROW_BOX) / 6, k = (row - ROW_BOX) % 6;
            return (k < WD && j == NV + WD * ci + k) ? 1.0 : 0.0;
        }
        if (CONES && row < ROW_TAU) {
            const int ci = (row - ROW_CONE) / 5, jr = (row - ROW_CONE) % 5;
            const int k = j - (NV + WD * ci);
            if (k < 0 || k >= 3) return 0.0;
            const double* R = g + OFF_CONE + 10 * ci;
            const double mu = R[9] * 0.70710678118654752440;
            const double c0 = jr == 0 ? 1.0 : (jr == 1 ? -1.0 : 0.0);
            const double c1 = jr == 2 ? 1.0 : (jr == 3 ? -1.0 : 0.0);
            const double c2 = jr == 4 ? -1.0 : -mu;
            return c0 * R[3 * k] + c1 * R[3 * k + 1] + c2 * R[3 * k + 2];
        }
        return j < NV ? g[OFF_JW + (row - ROW_OPT) * NV + j] : 0.0;
    }
    // dense task rows of a level (coefficient, right-hand side) and the diagonal (postural) part
    __device__ static __forceinline__ int task_rows(int level) { return level == 0 ? MD0 : MD1; }
    __device__ static __forceinline__ double task_coef(const double* g, int level, int r, int j, const Params& prm)
    {
        return j < NV ? (level == 0 ? prm.sw[0] : prm.sw[2]) * g[(level == 0 ? OFF_JW : OFF_JC) + r * NV + j] : 0.0;
    }
    __device__ static __forceinline__ double task_rhs(const double* g, int level, int r, const Params& prm)
    {
        const int t = level == 0 ? r : 6 + r;
        return (level == 0 ? prm.sw[0] : prm.sw[2]) * prm.lam * (g[OFF_RHS + t] - g[OFF_JDQD + t]);
    }
    __device__ static __forceinline__ void task_diag(const double* g, int level, int j, double& dgv, double& dbv, const Params& prm)
    {
        const bool post = level == 1 && j < NV && !(prm.post_act_only && j < 6);
        dgv = post ? prm.sw[1] * prm.sw[1] : 0.0; dbv = post ? prm.lam * g[OFF_RHS + 6 * (1 + NC) + j] : 0.0;
    }

    // row id and bounds of a dense slot (q >= NI_CHEAP)
    __device__ static __forceinline__ void dense_slot_bounds(const double* rec, int q, int& row, double& lo, double& hi)
    {
        const int a = q - NI_FORCE;
        row = ROW_TAU + a;
        const double ha = rec[OFF_H - SB + 6 + a];
        lo = rec[OFF_TAULIM - SB + a] - ha; hi = rec[OFF_TAULIM - SB + NA + a] - ha;
    }

    // Level-0 task value A0 x0* (6 numbers) -> eopt.
    template <int TEAM>
    __device__ static void task0_value(const double*, const double* g, const double* x, double* eopt, int tid)
    {
        if (tid < QPPVM_M0) {
            double s0 = 0.0, s1 = 0.0;
            const double* Jr = g + OFF_JW + tid * NV;
            int j = 0;
            for (; j + 1 < NV; j += 2) { s0 = fma(Jr[j], x[j], s0); s1 = fma(Jr[j + 1], x[j + 1], s1); }
            if (j < NV) s0 = fma(Jr[j], x[j], s0);
            eopt[tid] = s0 + s1;
        }
    }

    // tau = (M qdd + h - sum J_c^T [f;0]) actuated rows  (ref:src/ForceAcc.cpp:206-219)
    // (tv: the values M_a qdd - J_a^T f left by the last scan of the torque-limit slots, shapes with TAUVAL)
    template <int TEAM>
    __device__ static void recover(const double* rec, const double* g, const double* x, const double* tv, double* tau_out, bool ok, int tid)
    {
        for (int a = tid; a < NA; a += TEAM) {
            double v = 0.0;
            if (ok && TAUVAL) v = rec[OFF_H - SB + 6 + a] + tv[a];
            else if (ok) {
                const int i = 6 + a;
                v = rec[OFF_H - SB + i];
                for (int j = 0; j < NV; ++j) v = fma(M(rec, i, j), x[j], v);
                for (int j = 0; j < WD * NC; ++j) v = fma(-jcl(rec, g, j / WD, j % WD, i), x[NV + j], v);
            }
            tau_out[a] = v;     // failure: nothing is commanded (ForceAcc.cpp:189-193) -> zeros
        }
    }
};

// ------------------------------------------------------------------------------------------
// Problem policy: Torque stack  x = tau (fixed base)   (ref:src/QPPVMPlugin.cpp:112-188, 201-259)
//   level 0: (ee_right + ee_left): A = (J M^-1)[rows 0..2], b = A J^T F, F = K e + D edot   
//   level 1: joint impedance:       A = M^-1, b = M^-1 (K (q_ref - q) + D (-qdot))          
//   bounds : tau_min_const - h <= tau <= tau_max_const - h                                  (cpp:203-205)
//   output : tau_d = tau_qp + h, and tau_qp = 0 when the solve fails                        (cpp:246-256)
// Policy scratch (ext): M^-1 (N x LDM, symmetric) | A0 (6 x N, kept for the level-1 optimality rows) | T (N x LDM).
// ------------------------------------------------------------------------------------------
//   QPPVM_FLAG_JOINT_LIMITS: torque-domain JointLimits (cpp:169-171) -- simple bounds on the same variable, intersected
//                            with the shifted torque limits (AutoStack::getBounds)
//   QPPVM_FLAG_ELBOW_TASKS : level 1 = elbow_left + elbow_right (cpp:154-166; the alternative stack of cpp:177-178),
//                            A = (J_elbow M^-1)[rows 0..2], b = A J^T F; SEMIDEF like the hands -> regularised
template <int NA_, int FLAGS_ = 0>
struct Torque {
    static constexpr int KIND = QPPVM_KIND_TORQUE;
    static constexpr int NA = NA_, NC = 2, FLAGS = FLAGS_;
    static constexpr bool JLIM = (FLAGS & QPPVM_FLAG_JOINT_LIMITS) != 0, ELBOW = (FLAGS & QPPVM_FLAG_ELBOW_TASKS) != 0;
    static constexpr int NV = NA, N = NA, NB = NA;
    static constexpr int MD0 = 6, MD1 = ELBOW ? 6 : NA, MD_MAX = MD1;
    static constexpr int ROW_BOX = 0, ROW_OPT = NA, NROWS = NA + QPPVM_M0;
    static constexpr int WD = 3;
    static constexpr int NI = NA, NI_CHEAP = 0;              // every bound row is a (dense) row of J in whitened coordinates
    static constexpr bool TAUVAL = false;
    static constexpr int NTV = 0, NCT = 0;
    __device__ static __forceinline__ void dense_slot_bounds(const double*, int, int&, double&, double&) {}
    template <int TEAM> __device__ static __forceinline__ void fill_slot_table(const double*, double*, int) {}
    __device__ static __forceinline__ bool sparse_row(const double*, int, int&, double&, double&, double&) { return false; }
    static constexpr int OFF_J = 0, OFF_M = 12 * NA, OFF_H = OFF_M + NA * (NA + 1) / 2, OFF_FEE = OFF_H + NA;
    static constexpr int OFF_TAUJ = OFF_FEE + 12, OFF_TAULIM = OFF_TAUJ + NA;
    static constexpr int OFF_JLIM = OFF_TAULIM + 2 * NA, OFF_JEL = OFF_JLIM + (JLIM ? 2 * NA : 0), OFF_FEL = OFF_JEL + 12 * NA;
    static constexpr int REC_UNPADDED = OFF_JEL + (ELBOW ? 12 * NA + 12 : 0);
    static constexpr int REC = REC_UNPADDED + (REC_UNPADDED & 1);
    static constexpr int LDM = NA | 1;
    static constexpr bool STAGE_RECORD = false;                // see Slab::STAGE
    static constexpr bool EXT_IS_GLOBAL = false;               // `ext` is the policy scratch (M^-1, A0, T)
    static constexpr int STAGE_FROM = 0, STAGED = 0, S_JCL = 0;
    static constexpr int KMAX = kmax_for(6, NA, N);
    static constexpr int O_A0 = NA * LDM, O_T = O_A0 + 6 * NA;
    static constexpr int EXTRA = O_T + NA * LDM + ((O_T + NA * LDM) & 1);
    static constexpr bool SPLIT_FACTOR = false;
    static constexpr bool J_GLOBAL = false;
    static constexpr bool HAS_EQ_COEF = false;

    __device__ static __forceinline__ int n_eq(int level) { return level == 0 ? 0 : 6; }
    __device__ static __forceinline__ int eq_row(int, int e) { return ROW_OPT + e; }
    // CartesianImpedanceCtrl: HST_SEMIDEF -> regularised; JointImpedanceCtrl: HST_POSDEF -> not
    __device__ static __forceinline__ bool regularised(int level) { return level == 0 || ELBOW; }
    // bounds of tau_i: shifted torque limits (cpp:203-205), intersected with the joint-limit bounds when present
    __device__ static __forceinline__ void bounds(const double* rec, int i, double& lo, double& hi)
    {
        const double h = rec[OFF_H + i];
        lo = rec[OFF_TAULIM + i] - h; hi = rec[OFF_TAULIM + NA + i] - h;
        if (JLIM) { lo = fmax(lo, rec[OFF_JLIM + i]); hi = fmin(hi, rec[OFF_JLIM + NA + i]); }
    }

    // M^-1 by Cholesky (M = L L^T), T = L^-1, M^-1 = T^T T; then A0 = (J_t M^-1)[0..2] for both hands.
    // Element (i, j) of a square buffer X lives at X[j * LDM + i].
    template <int TEAM>
    __device__ static bool prepare(const double* rec, double* ext, int tid)
    {
        double* const L = ext;
        double* const A0 = ext + O_A0;
        double* const T = ext + O_T;
        for (int t = tid; t < NA * NA; t += TEAM) {
            const int i = t / NA, j = t - i * NA;
            L[j * LDM + i] = i >= j ? rec[OFF_M + i * (i + 1) / 2 + j] : 0.0;
            T[j * LDM + i] = 0.0;
        }
        Team<TEAM>::sync();
        bool ok = true;
#pragma unroll 1
        for (int k = 0; k < NA; ++k) {                       // right-looking Cholesky, thread i owns row i
            const double d = L[k * LDM + k];
            if (!(d > 0.0)) ok = false;                      // team-uniform
            const double inv = rsqrt(d);
            Team<TEAM>::sync();
            for (int i = k + tid; i < NA; i += TEAM) L[k * LDM + i] *= inv;       // column k (diagonal: d / sqrt(d))
            Team<TEAM>::sync();
            for (int i = k + 1 + tid; i < NA; i += TEAM) {
                const double lik = L[k * LDM + i];
#pragma unroll 2
                for (int j = k + 1; j <= i; ++j) L[j * LDM + i] = fma(-lik, L[k * LDM + j], L[j * LDM + i]);
            }
            Team<TEAM>::sync();
        }
        if (!ok) return false;
        for (int c = tid; c < NA; c += TEAM) {               // T = L^-1 (lower), thread c owns column c
            T[c * LDM + c] = 1.0 / L[c * LDM + c];
#pragma unroll 1
            for (int i = c + 1; i < NA; ++i) {
                double sacc = 0.0;
#pragma unroll 2
                for (int l = c; l < i; ++l) sacc = fma(L[l * LDM + i], T[c * LDM + l], sacc);
                T[c * LDM + i] = -sacc / L[i * LDM + i];
            }
        }
        Team<TEAM>::sync();
        for (int t = tid; t < NA * NA; t += TEAM) {          // M^-1 = T^T T over the dead L
            const int i = t / NA, j = t - i * NA;
            double sacc = 0.0;
#pragma unroll 2
            for (int k = (i > j ? i : j); k < NA; ++k) sacc = fma(T[i * LDM + k], T[j * LDM + k], sacc);
            L[j * LDM + i] = sacc;
        }
        Team<TEAM>::sync();
        for (int t = tid; t < 6 * NA; t += TEAM) {           // A0 row (3 t + r) = J_t[r] M^-1
            const int row = t / NA, j = t - row * NA;
            const double* Jr = rec + OFF_J + ((row / 3) * 6 + (row % 3)) * NA;
            double sacc = 0.0;
#pragma unroll 2
            for (int k = 0; k < NA; ++k) sacc = fma(Jr[k], L[j * LDM + k], sacc);
            A0[row * NA + j] = sacc;
        }
        if (ELBOW) {
            // A1 row (3 t + r) = J_elbow_t[r] M^-1 over the dead T (the products above no longer read it)
            for (int t = tid; t < 6 * NA; t += TEAM) {
                const int row = t / NA, j = t - row * NA;
                const double* Jr = rec + OFF_JEL + ((row / 3) * 6 + (row % 3)) * NA;
                double sacc = 0.0;
#pragma unroll 2
                for (int k = 0; k < NA; ++k) sacc = fma(Jr[k], L[j * LDM + k], sacc);
                T[row * NA + j] = sacc;
            }
        }
        Team<TEAM>::sync();
        return true;
    }
    // b = A (J^T F) for the 3-row Cartesian impedance tasks (full 6-row J and 6-vector F): threads 0 .. 5
    __device__ static __forceinline__ double cart_rhs(const double* Arow, const double* J, const double* F)
    {
        double sacc = 0.0;
        for (int j = 0; j < NA; ++j) {
            double jtf = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) jtf = fma(J[r * NA + j], F[r], jtf);
            sacc = fma(Arow[j], jtf, sacc);
        }
        return sacc;
    }

    template <int TEAM>
    __device__ static int load_tasks(const double* rec, const double* ext, int level, double* Ad, double* dg, double* db, int tid,
                                     const Params&)
    {
        constexpr int LDA = NB + 1;
        const double* Minv = ext;
        const double* A0 = ext + O_A0;
        for (int j = tid; j < N; j += TEAM) { dg[j] = 0.0; db[j] = 0.0; }
        if (level == 0) {
            for (int t = tid; t < 6 * NA; t += TEAM) { const int r = t / NA, j = t - r * NA; Ad[r * LDA + j] = A0[t]; }
            if (tid < 6) Ad[tid * LDA + NB] = cart_rhs(A0 + tid * NA, rec + OFF_J + (tid / 3) * 6 * NA, rec + OFF_FEE + 6 * (tid / 3));
            return 6;
        }
        if constexpr (ELBOW) {
            const double* A1 = ext + O_T;
            for (int t = tid; t < 6 * NA; t += TEAM) { const int r = t / NA, j = t - r * NA; Ad[r * LDA + j] = A1[t]; }
            if (tid < 6) Ad[tid * LDA + NB] = cart_rhs(A1 + tid * NA, rec + OFF_JEL + (tid / 3) * 6 * NA, rec + OFF_FEL + 6 * (tid / 3));
            return 6;
        } else {
            for (int t = tid; t < NA * NA; t += TEAM) { const int r = t / NA, j = t - r * NA; Ad[r * LDA + j] = Minv[j * LDM + r]; }
            for (int r = tid; r < NA; r += TEAM) {           // b = M^-1 tau_j
                double sacc = 0.0;
#pragma unroll 2
                for (int j = 0; j < NA; ++j) sacc = fma(Minv[j * LDM + r], rec[OFF_TAUJ + j], sacc);
                Ad[r * LDA + NB] = sacc;
            }
            return NA;
        }
    }

    template <int TEAM>
    __device__ static void build_row(const double* rec, const double* ext, int row, const double* eopt,
                                     double* av, double& lo, double& hi, int tid)
    {
        if (row < ROW_OPT) {                                 // TorqueLimits: simple bound on tau_row
            for (int j = tid; j < N; j += TEAM) av[j] = j == row ? 1.0 : 0.0;
            bounds(rec, row, lo, hi);
        } else {                                             // optimality rows: A0 x = A0 x0*
            const int r = row - ROW_OPT;
            for (int j = tid; j < N; j += TEAM) av[j] = ext[O_A0 + r * NA + j];
            lo = hi = eopt[r];
        }
    }
    __device__ static void eval_slot(const double* rec, const double*, const double*, int q, const double* x,
                                     int& row, double& val, double& lo, double& hi)
    {
        row = q; val = x[q];
        bounds(rec, q, lo, hi);
    }
    template <int TEAM>
    __device__ static void task0_value(const double*, const double* ext, const double* x, double* eopt, int tid)
    {
        if (tid < QPPVM_M0) {
            double sacc = 0.0;
            for (int j = 0; j < NA; ++j) sacc = fma(ext[O_A0 + tid * NA + j], x[j], sacc);
            eopt[tid] = sacc;
        }
    }
    // tau_d = tau_qp + h ; tau_qp = 0 on failure (ref:src/QPPVMPlugin.cpp:246-256)
    template <int TEAM>
    __device__ static void recover(const double* rec, const double*, const double* x, const double*, double* tau_out, bool ok, int tid)
    {
        for (int a = tid; a < NA; a += TEAM) tau_out[a] = (ok ? x[a] : 0.0) + rec[OFF_H + a];
    }
};

// ------------------------------------------------------------------------------------------
// Shared-memory slab of one team
// ------------------------------------------------------------------------------------------
template <class P>
struct Slab {
    static constexpr int N = P::N, NB = P::NB;
    static constexpr int KMAX = P::KMAX;
    static constexpr int LDQ = KMAX | 1;              // row stride of Q1 (odd: conflict-free 64-bit column walks)
    static constexpr int LDR = KMAX | 1;              // column stride of RN
    static constexpr int KP = (KMAX + 1) & ~1;        // even padding of the small per-constraint vectors
    static constexpr int LDJ = NB | 1;                // odd column stride
    static constexpr int LDA = NB + 1;
    static constexpr int VEC = (N + 3) & ~3;
    // offsets in doubles from the slab base
    static constexpr int O_REC = 0;                   // staged record (16-B aligned: first in the slab)
    // Stage the record in shared memory (one TMA bulk copy) or leave it in global memory behind L1/L2.  Measured
    // per shape (profiles/README.md): shapes whose inequality scan re-reads M every iteration (torque-limit rows)
    // want it staged; for the others the 11-18 KB are worth more as 4 extra resident CTAs per SM.
    static constexpr bool STAGE = P::STAGE_RECORD;
    static constexpr int O_J = O_REC + (STAGE ? P::STAGED : 0) + 4;   // staged part | pointer slots: J (global), record, workspace
    // J (and R before it) packed upper-triangular: NB (NB + 1) / 2 doubles.  R is row-major packed
    // (R(i,l) at i NB - i (i - 1) / 2 + (l - i)), J column-major packed (J(i,j) at j (j + 1) / 2 + i): triangular
    // numbers are a permutation mod 16, so a half-warp walking 16 consecutive columns is bank-conflict free.
    static constexpr int SZ_J = NB * (NB + 1) / 2 + ((NB * (NB + 1) / 2) & 1);
    static constexpr int O_Q = O_J + (P::J_GLOBAL ? 0 : SZ_J);   // Q1, aliased by Ad during the factorisation
    static constexpr int SZ_Q_RAW = (N * LDQ > P::MD_MAX * LDA) ? N * LDQ : P::MD_MAX * LDA;
    static constexpr int SZ_Q = SZ_Q_RAW + (SZ_Q_RAW & 1);     // even: RN and the vectors stay 16-byte aligned
    // RN (triangular factor of the whitened active normals) and its inverse RI, both packed by columns:
    // (r, c), r <= c, at c (c + 1) / 2 + r.  RI turns every "solve with RN" of the active-set iterations into a
    // thread-parallel product (a back substitution is k dependent steps on one warp while the other waits).
    static constexpr int SZ_TRI = KMAX * (KMAX + 1) / 2 + ((KMAX * (KMAX + 1) / 2) & 1);
    static constexpr int O_R = O_Q + SZ_Q;
    static constexpr int O_RI = O_R + SZ_TRI;
    static constexpr int O_VEC = O_RI + SZ_TRI;       // u0 u x w w2 av dg db xp jd
    // dg / db hold the diagonal task part for the in-kernel factorisation and KKT check; the shapes that do both in
    // other kernels only keep the part of dg that the 64-double exchange buffer of gs_dots spills into
    static constexpr int SZ_DG = P::SPLIT_FACTOR ? ((VEC < 64 ? 64 - VEC : 0) + 1) / 2 * 2 : VEC, SZ_DB = P::SPLIT_FACTOR ? 0 : VEC;
    static constexpr int O_DG = O_VEC + 6 * VEC, O_DB = O_DG + SZ_DG, O_XP = O_DB + SZ_DB, O_JD = O_XP + VEC;
    static constexpr int O_TV = O_JD + VEC;           // values of the dense inequality slots at the last full scan
    static constexpr int O_CT = O_TV + P::NTV;        // per force-only slot: three coefficients, lower, upper bound
    static constexpr int O_SMALL = O_CT + P::NCT;     // d1 rr lam (KP each) | eopt 8 | red 16
    static constexpr int O_MBAR = O_SMALL + 6 * KP + 8 + 16;   // d1 rr lam rdi gc gs | 2 mbarriers: record staging, workspace copies
    static constexpr int O_PRM = O_MBAR + 2;          // the kernel's Params (for the policy functions called inside the solver)
    static constexpr int O_STATE = O_PRM + (sizeof(Params) + 7) / 8;   // ints: k, n_act_ineq, iters, ws phase | act_row[KP] | act_sgn[KP]
    static constexpr int O_CSTATE = O_STATE + 2 + KP;     // bytes
    static constexpr int O_EXT = O_CSTATE + ((P::NROWS + 15) & ~15) / 8;   // policy scratch
    static constexpr int DOUBLES = O_EXT + P::EXTRA;
    // prepare workspace per level, every piece in the layout (and 16-byte alignment) of its shared-memory home so
    // that the solve kernel fetches it with bulk copies: J | u0 | jd | Q1 rows [i][LDQ] (first n_eq columns
    // filled) | RN columns [c][LDR] | 1/diag + fallback flag | point after the equalities with a known rhs
    static constexpr int NEQ_MAX = 12;
    static constexpr int WS_U0 = SZ_J, WS_JD = WS_U0 + VEC, WS_Q = WS_JD + VEC;
    // (staged fetch of the compact Q1: it lands in the last WSZ_Q doubles of the Q1 region)
    static constexpr int Q_STAGE = SZ_Q - (N * NEQ_MAX + ((N * NEQ_MAX) & 1));
    static_assert(!P::SPLIT_FACTOR || (Q_STAGE >= 0 && ((O_Q + Q_STAGE) % 2) == 0), "staging block inside the Q1 region, 16-byte aligned");
#if QPPVM_WS_COMPACT_Q
    // Q1 of the equality rows, compact: (i, c) at i n_eq + c -- 2.4 / 4.9 KB per level instead of the 12.6 KB block of the
    // shared-memory layout (whose columns >= n_eq the prepare kernel never wrote, but the bulk copy fetched)
    static constexpr int WSZ_Q = N * NEQ_MAX + ((N * NEQ_MAX) & 1), WS_RN = WS_Q + WSZ_Q;
#else
    static constexpr int WSZ_Q = N * LDQ + ((N * LDQ) & 1), WS_RN = WS_Q + WSZ_Q;
#endif
    static constexpr int WSZ_RN = NEQ_MAX * (NEQ_MAX + 1) / 2 + ((NEQ_MAX * (NEQ_MAX + 1) / 2) & 1), WS_RDI = WS_RN + WSZ_RN;
    static constexpr int WSZ_RDI = NEQ_MAX + 2, WS_FLAG = WS_RDI + NEQ_MAX, WS_U = WS_RDI + WSZ_RDI;
    static constexpr int WS_RI = WS_U + VEC;          // inverse of the equality block of RN, packed like RN
    static constexpr int WS_LEVEL = WS_RI + WSZ_RN;
    static_assert((SZ_J % 2 == 0) && (VEC % 2 == 0) && (WSZ_RDI % 2 == 0) && (KP >= WSZ_RDI), "bulk-copy alignment");
    static constexpr int WS = 2 * WS_LEVEL;
    // Certificate block: what qp_certify_kernel needs from the solve kernel, written over the (consumed) head of the
    // level's workspace block: x | proximal centre | level-0 task value | signed multipliers | k, rows, signs (ints)
    static constexpr int C_X = 0, C_XP = VEC, C_EOPT = 2 * VEC, C_Y = 2 * VEC + 8, C_INT = C_Y + KP, C_END = C_INT + KP + 2;
    static_assert(C_END <= WS_LEVEL, "certificate block fits in the level's workspace block");
    static constexpr int BYTES = DOUBLES * 8;
    // resident CTAs per SM the slab allows (228 KB of shared memory, 1 KB reserved per CTA), capped at 10: beyond
    // that the throughput was flat (profiles/README.md); __launch_bounds__ turns it into the register budget
    static constexpr int CTAS_RAW = 233472 / (BYTES + 1024);
    static constexpr int CTAS = CTAS_RAW > 10 ? 10 : (CTAS_RAW < 1 ? 1 : CTAS_RAW);
};

// Sums NV per-lane values across the warp with NV - 1 + (5 - log2 NV) shuffles instead of 5 NV: at each of the
// first log2 NV steps a lane keeps the half of its values selected by one lane-id bit and sends the other half.
// Returns, in every lane l, the full sum of value l >> (5 - log2 NV).
template <int NV>
__device__ __forceinline__ double warp_sum_transposed(double (&p)[NV], int l)
{
    int off = 16;
#pragma unroll
    for (int half = NV / 2; half >= 1; half /= 2, off /= 2) {
        const bool up = l & off;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double send = up ? p[i] : p[i + half], keep = up ? p[i + half] : p[i];
            p[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    double v = p[0];
#pragma unroll
    for (; off >= 1; off /= 2) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ------------------------------------------------------------------------------------------
// Whitening factorisation, shared by the in-kernel path (one 64-thread team, lane = thread id) and the
// lane-batched factor kernel (several factorisations side by side in one CTA, GS = NB + 1 lanes each).
// R^T R = D + eps I + Ad^T Ad via Householder QR of the stacked matrix [sqrt(D+eps); Ad], then J = R^-1, on the
// leading NB x NB block; columns >= NB carry no dense task entries and stay diagonal (jd).
// u0 = R^-T (D db + Ad^T b) comes out as the transformed right-hand side (the proximal centre is zero whenever a
// level is factorised).  Register blocking: lane j keeps column j of Ad (MD doubles; j == NB is the rhs column) in
// registers for the whole QR, the pivot column is broadcast through a double-buffered smem vector (one barrier
// per step); then lane j keeps column j of J in registers through a fully unrolled back substitution.
// Every thread of the CTA must call it (barriers); lanes with j > NB only take part in the barriers and in the
// strided loops, a group that has nothing to do passes j = -1.
// ------------------------------------------------------------------------------------------
template <int MD, int N, int NB, int GS>
__device__ __noinline__ void factor_qr(int j, double* Jm, const double* Ad, const double* dg, const double* db,
                                          double* u0, double* jd, double* bc, double eps)
{
    constexpr int LDA = NB + 1;
    constexpr int BC = MD + 4;
    const bool live = j >= 0;
    double* const rinv = jd;               // 1 / R(i,i) for i < NB (jd proper only uses i >= NB)
    if (live) for (int i = j; i < NB * (NB + 1) / 2; i += GS) Jm[i] = 0.0;
    double col[MD];
#pragma unroll
    for (int r = 0; r < MD; ++r) col[r] = (live && j <= NB) ? Ad[r * LDA + j] : 0.0;
    if (live)
        for (int i = j; i < N; i += GS) {
            const double dd = dg[i] + eps;
            const double rt = sqrt(dd);
            if (i < NB) rinv[i] = 1.0 / rt; else jd[i] = 1.0 / rt;
            u0[i] = dd > 0.0 ? (dg[i] * db[i]) / rt : 0.0;
        }
    __syncthreads();
    double top = 0.0;                      // R(kc, j) before the step is zero except for the rhs column (u0[kc])
#pragma unroll 1
    for (int kc = 0; kc < NB; ++kc) {
        double* const b = bc + (kc & 1) * BC;
        if (j == kc) {                     // pivot owner: reflector of [alpha; col]
            double sg0 = 0.0, sg1 = 0.0;
#pragma unroll
            for (int r = 0; r + 1 < MD; r += 2) { sg0 = fma(col[r], col[r], sg0); sg1 = fma(col[r + 1], col[r + 1], sg1); }
            if (MD & 1) sg0 = fma(col[MD - 1], col[MD - 1], sg0);
            const double sigma = sg0 + sg1;
            double v1 = 0.0, tau = 0.0;
            const double alpha = sqrt(dg[kc] + eps);
            if (sigma != 0.0) {
                const double nrm = sqrt(fma(alpha, alpha, sigma));
                v1 = -sigma / (alpha + nrm);               // alpha - nrm, cancellation-free (alpha >= 0)
                tau = 2.0 / fma(v1, v1, sigma);
                rinv[kc] = 1.0 / nrm;
            }
#pragma unroll
            for (int r = 0; r < MD; ++r) b[r] = col[r];
            b[MD] = v1; b[MD + 1] = tau;
        }
        __syncthreads();
        if (live) {
            const double tau = b[MD + 1];
            if (tau != 0.0 && j > kc && j <= NB) {             // tau == 0: empty pivot column, nothing to do
                const double v1 = b[MD];
                top = j == NB ? u0[kc] : 0.0;
                double s0 = v1 * top, s1 = 0.0;
#pragma unroll
                for (int r = 0; r + 1 < MD; r += 2) { s0 = fma(b[r], col[r], s0); s1 = fma(b[r + 1], col[r + 1], s1); }
                if (MD & 1) s0 = fma(b[MD - 1], col[MD - 1], s0);
                const double sc = (s0 + s1) * tau;
                if (j == NB) u0[kc] = top - sc * v1; else Jm[kc * NB - kc * (kc - 1) / 2 + (j - kc)] = -sc * v1;   // R(kc, j)
#pragma unroll
                for (int r = 0; r < MD; ++r) col[r] = fma(-sc, b[r], col[r]);
            }
        }
    }
    __syncthreads();
}
// Second half (independent of the number of dense task rows: one copy of the unrolled code serves both levels):
// back substitution, lane j holds column j of J: J(i,j) = (d_ij - sum_{l>i} R(i,l) J(l,j)) / R(i,i).
// Uniform over lanes: Jc[l] stays 0 for l > j.  Row i of R is a broadcast read.
template <int NB>
__device__ __noinline__ void factor_backsub(int j, double* Jm, const double* jd)
{
    const bool live = j >= 0;
    const double* const rinv = jd;
    double Jc[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) Jc[i] = 0.0;
    // (lanes without a column run it too, on zeros; predicates around these unrolled rows have made ptxas demote Jc
    // to local memory: profiles/README.md, skipped-back-substitution experiment)
#pragma unroll
    for (int i = NB - 1; i >= 0; --i) {
        double a0 = (j == i) ? 1.0 : 0.0, a1 = 0.0;
#pragma unroll
        for (int l = i + 1; l < NB; ++l) {
            const double ril = Jm[i * NB - i * (i - 1) / 2 + (l - i)];
            if ((l - i) & 1) a0 = fma(-ril, Jc[l], a0); else a1 = fma(-ril, Jc[l], a1);
        }
        Jc[i] = (a0 + a1) * rinv[i];
    }
    __syncthreads();                       // every row of R has been consumed: overwrite with J (packed columns)
    if (live && j < NB) {
#pragma unroll
        for (int i = 0; i < NB; ++i) if (i <= j) Jm[j * (j + 1) / 2 + i] = Jc[i];
    }
    __syncthreads();
}
template <int MD, int N, int NB, int GS>
__device__ __forceinline__ void factor_core(int j, double* Jm, const double* Ad, const double* dg, const double* db,
                                            double* u0, double* jd, double* bc, double eps)
{
    factor_qr<MD, N, NB, GS>(j, Jm, Ad, dg, db, u0, jd, bc, eps);
    factor_backsub<NB>(j, Jm, jd);
}

// ------------------------------------------------------------------------------------------
// The solver
// ------------------------------------------------------------------------------------------
template <class P, int TEAM>
struct Solver {
    using S = Slab<P>;
    static constexpr int N = P::N, NB = P::NB, LDJ = S::LDJ, LDA = S::LDA;
    static constexpr int KMAX = S::KMAX, LDQ = S::LDQ, LDR = S::LDR, KP = S::KP;

#define QP_SM(name, off) __device__ static __forceinline__ double* name##_() { return reinterpret_cast<double*>(g_smem) + (off); }
    __device__ static __forceinline__ double* grec_()         // the record in global memory
    {
        return *reinterpret_cast<double**>(reinterpret_cast<double*>(g_smem) + S::O_J - 2);
    }
    __device__ static __forceinline__ double*& ws_()           // this problem's factor workspace (global memory)
    {
        return *reinterpret_cast<double**>(reinterpret_cast<double*>(g_smem) + S::O_J - 1);
    }
    __device__ static __forceinline__ double* rec_()
    {
        if (S::STAGE) return reinterpret_cast<double*>(g_smem) + S::O_REC;
        return grec_();
    }
    QP_SM(Q1, S::O_Q) QP_SM(Ad, S::O_Q) QP_SM(RN, S::O_R)
    __device__ static __forceinline__ double*& jptr_()          // J of the current level in global memory (J_GLOBAL shapes)
    {
        return *reinterpret_cast<double**>(reinterpret_cast<double*>(g_smem) + S::O_J - 3);
    }
    __device__ static __forceinline__ double* Jm_()
    {
        if (P::J_GLOBAL) return jptr_();
        return reinterpret_cast<double*>(g_smem) + S::O_J;
    }
    QP_SM(u0, S::O_VEC) QP_SM(u, S::O_VEC + S::VEC) QP_SM(x, S::O_VEC + 2 * S::VEC) QP_SM(w, S::O_VEC + 3 * S::VEC)
    QP_SM(w2, S::O_VEC + 4 * S::VEC) QP_SM(av, S::O_VEC + 5 * S::VEC) QP_SM(dg, S::O_DG)
    QP_SM(db, S::O_DB) QP_SM(xp, S::O_XP) QP_SM(jd, S::O_JD)
    QP_SM(d1, S::O_SMALL) QP_SM(rr, S::O_SMALL + KP) QP_SM(lam, S::O_SMALL + 2 * KP) QP_SM(rdi, S::O_SMALL + 3 * KP)
    QP_SM(gc, S::O_SMALL + 4 * KP) QP_SM(gsn, S::O_SMALL + 5 * KP)
    QP_SM(eopt, S::O_SMALL + 6 * KP) QP_SM(red, S::O_SMALL + 6 * KP + 8) QP_SM(ext, S::O_EXT) QP_SM(tv, S::O_TV)
    QP_SM(RI, S::O_RI) QP_SM(ct, S::O_CT)
    __device__ static __forceinline__ constexpr int tri(int c) { return c * (c + 1) / 2; }   // start of packed column c
#undef QP_SM
    __device__ static __forceinline__ Params* prm_() { return reinterpret_cast<Params*>(reinterpret_cast<double*>(g_smem) + S::O_PRM); }
    __device__ static __forceinline__ uint64_t* mbar_() { return reinterpret_cast<uint64_t*>(g_smem) + S::O_MBAR; }
    __device__ static __forceinline__ uint64_t* mbar_ws_() { return reinterpret_cast<uint64_t*>(g_smem) + S::O_MBAR + 1; }
    __device__ static __forceinline__ int* state_() { return reinterpret_cast<int*>(reinterpret_cast<double*>(g_smem) + S::O_STATE); }
    __device__ static __forceinline__ unsigned char* cstate_() { return g_smem + 8 * S::O_CSTATE; }
    using tm = Team<TEAM>;
// Local aliases of the slab regions (constant addresses: no registers, no memory loads).
#define QP_BIND                                                                                         \
    double* const rec = rec_(); double* const Jm = Jm_(); double* const Q1 = Q1_(); double* const Ad = Ad_(); \
    double* const RN = RN_(); double* const u0 = u0_(); double* const u = u_(); double* const x = x_();     \
    double* const w = w_(); double* const w2 = w2_(); double* const av = av_(); double* const dg = dg_();   \
    double* const db = db_(); double* const xp = xp_(); double* const jd = jd_(); double* const d1 = d1_(); \
    double* const rr = rr_(); double* const lam = lam_(); double* const eopt = eopt_(); double* const red = red_(); \
    double* const ext = P::EXT_IS_GLOBAL ? grec_() : ext_(); (void)ext; double* const rdi = rdi_(); (void)rdi; \
    double* const RI = RI_(); (void)RI; double* const ct = ct_(); (void)ct;                                    \
    double* const gc = gc_(); (void)gc; double* const gsn = gsn_(); (void)gsn;                                 \
    int* const st = state_(); int* const act_row = st + 4; int* const act_sgn = st + 4 + KP;               \
    unsigned char* const cstate = cstate_(); const int tid = threadIdx.x;                                  \
    (void)rec; (void)Jm; (void)Q1; (void)Ad; (void)RN; (void)u0; (void)u; (void)x; (void)w; (void)w2; (void)av; \
    (void)dg; (void)db; (void)xp; (void)jd; (void)d1; (void)rr; (void)lam; (void)eopt; (void)red; (void)st; \
    (void)act_row; (void)act_sgn; (void)cstate; (void)tid;
    // st[0] = k (active constraints), st[1] = active inequalities, st[2] = working-set changes

    // ---- whitening: R^T R = D + eps I + Ad^T Ad via Householder QR of the stacked matrix [sqrt(D+eps); Ad],
    // then J = R^-1, on the leading NB x NB block; columns >= NB carry no dense task entries and stay
    // diagonal (jd).  u0 = R^-T (D db + Ad^T b + eps xp) comes out as the transformed right-hand side.
    // Register blocking: thread j keeps column j of Ad (MD doubles; j == NB is the rhs column) in registers for
    // the whole QR, the pivot column is broadcast through a double-buffered smem vector (one barrier per
    // step); then thread j keeps column j of J in registers through a fully unrolled back substitution.
    template <int MD>
    __device__ static __noinline__ void factor(double eps)
    {
        static_assert(NB + 1 <= TEAM, "one task column per thread");
        static_assert(2 * (MD + 4) <= 3 * S::VEC, "broadcast buffer fits in w|w2|av");
        QP_BIND
        // w, w2, av are contiguous and dead here: 2 x (MD + 4) broadcast slots
        factor_core<MD, N, NB, TEAM>(tid, Jm, Ad, dg, db, u0, jd, w, eps);
    }

    // Triangular mat-vecs with J.  A warp executes as many iterations as its longest lane, and with one column
    // (row) per thread the long ones sit next to idle lanes: the part of a column beyond HALF entries goes to a
    // helper thread (tid >= NB), whose partial sum travels through d1 (scratch of gs_pass, free here).
    // (at most TEAM - NB helper lanes exist: wide shapes leave more of each column to the main lane)
    static constexpr int HALF = (NB - 1) / 2 > 2 * NB - 1 - TEAM ? (NB - 1) / 2 : 2 * NB - 1 - TEAM;   // main lanes: entries [0, HALF] of their column
    static constexpr int NHELP = NB - 1 - HALF;                // columns HALF + 1 .. NB - 1 have a helper
    static_assert(NB + NHELP <= TEAM && NHELP <= KP, "helper lanes and their exchange slots");

    // out = sgn * J^T a   (column j of J is contiguous in i)
    __device__ static __noinline__ void whiten(const double* a, double sgn, double* out)
    {
        QP_BIND
        double r = 0.0;
        if (tid < NB + NHELP) {
            const bool helper = tid >= NB;
            const int jc = helper ? HALF + 1 + (tid - NB) : tid;
            const double* col = Jm + jc * (jc + 1) / 2;
            int i = helper ? HALF + 1 : 0;
            const int last = helper ? jc : (jc < HALF ? jc : HALF);
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 2
            for (; i + 1 <= last; i += 2) { s0 = fma(col[i], a[i], s0); s1 = fma(col[i + 1], a[i + 1], s1); }
            if (i <= last) s0 = fma(col[i], a[i], s0);
            r = s0 + s1;
            if (helper) d1[tid - NB] = r;
        }
        tm::sync();
        if (tid < NB) out[tid] = sgn * (tid > HALF ? r + d1[tid - HALF - 1] : r);
        else if (tid < N) out[tid] = sgn * jd[tid] * a[tid];
        for (int jj = TEAM + tid; jj < N; jj += TEAM) out[jj] = sgn * jd[jj] * a[jj];
        tm::sync();
    }
    // xx = J uu  (row i of J: entries J(i, j), j >= i, at Jm[j (j + 1) / 2 + i]); main lanes take the last
    // HALF + 1 columns of their row, the helper of row i < NHELP the columns before those
    __device__ static __noinline__ void unwhiten(const double* uu, double* xx)
    {
        QP_BIND
        double r = 0.0;
        if (tid < NB + NHELP) {
            const bool helper = tid >= NB;
            const int ir = helper ? tid - NB : tid;
            int jx = helper ? ir : (ir > NHELP ? ir : NHELP);
            const int last = helper ? NHELP - 1 : NB - 1;
            const double* Ji = Jm + ir;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 2
            for (; jx + 1 <= last; jx += 2) {
                s0 = fma(Ji[jx * (jx + 1) / 2], uu[jx], s0);
                s1 = fma(Ji[(jx + 1) * (jx + 2) / 2], uu[jx + 1], s1);
            }
            if (jx <= last) s0 = fma(Ji[jx * (jx + 1) / 2], uu[jx], s0);
            r = s0 + s1;
            if (helper) d1[ir] = r;
        }
        tm::sync();
        if (tid < NB) xx[tid] = tid < NHELP ? r + d1[tid] : r;
        else if (tid < N) xx[tid] = jd[tid] * uu[tid];
        for (int jj = TEAM + tid; jj < N; jj += TEAM) xx[jj] = jd[jj] * uu[jj];
        tm::sync();
    }
    __device__ static __forceinline__ double dot(const double* a, const double* b)
    {
        double s = 0.0;
        for (int i = threadIdx.x; i < N; i += TEAM) s = fma(a[i], b[i], s);
        return tm::sum(s, red_());
    }

    // One Gram-Schmidt pass against the k active normals, in two halves: gs_dots: rr = Q1^T v, d1 (+)= rr;
    // gs_update: dst = src - Q1 rr.
    // The k dot products of length N are shared by all threads: column c is split over SEG = TEAM / CW threads
    // (CW = 16 or 32 columns wide), each walking every SEG-th row; the partial sums meet through a shuffle inside
    // the warp and a small exchange buffer across warps (`part`: WARPS x 32 doubles over av | dg, both dead during
    // the active-set iterations: dg / db are reloaded by kkt()).
    static_assert(S::VEC + S::SZ_DG >= Team<TEAM>::WARPS * 32, "the exchange buffer of gs_dots covers av | dg");
    __device__ static __noinline__ void gs_dots(const double* v, bool accumulate, int k)
    {
        QP_BIND
        double* const part = av;
        {
            const bool narrow = k <= 16;
            const int cw = narrow ? 16 : 32;
            const int c = tid & (cw - 1), seg = tid / cw, nseg = TEAM / cw;
            double s0 = 0.0, s1 = 0.0;
            if (c < k) {
                int i = seg;
#pragma unroll 2
                for (; i + nseg < N; i += 2 * nseg) {
                    s0 = fma(Q1[i * LDQ + c], v[i], s0);
                    s1 = fma(Q1[(i + nseg) * LDQ + c], v[i + nseg], s1);
                }
                if (i < N) s0 = fma(Q1[i * LDQ + c], v[i], s0);
            }
            double sacc = s0 + s1;
            if (narrow) sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);     // the two segments living in this warp
            if ((tid & 31) < cw && (tid & 31) == c) part[(tid >> 5) * 32 + c] = sacc;
        }
        tm::sync();
        if (tid < k) {
            double sacc = part[tid];
#pragma unroll
            for (int wv = 1; wv < tm::WARPS; ++wv) sacc += part[wv * 32 + tid];
            rr[tid] = sacc;                                  // rr: scratch for this pass' coefficients
            d1[tid] = accumulate ? d1[tid] + sacc : sacc;
        }
        tm::sync();
    }
    // The same for a normal with three entries (columns j0 .. j0 + 2: a force-only row): three products per column.
    __device__ static __forceinline__ void gs_dots_sparse(const double* v, int j0, int k)
    {
        QP_BIND
        if (tid < k) {
            const double sacc = fma(Q1[j0 * LDQ + tid], v[j0], fma(Q1[(j0 + 1) * LDQ + tid], v[j0 + 1], Q1[(j0 + 2) * LDQ + tid] * v[j0 + 2]));
            rr[tid] = sacc; d1[tid] = sacc;
        }
        tm::sync();
    }
    __device__ static __noinline__ void gs_update(const double* src, double* dst, int k)
    {
        QP_BIND
        for (int i = tid; i < N; i += TEAM) {
            double s0 = src[i], s1 = 0.0;
            const double* q = Q1 + i * LDQ;
            int c = 0;
#pragma unroll 2
            for (; c + 1 < k; c += 2) { s0 = fma(-q[c], rr[c], s0); s1 = fma(-q[c + 1], rr[c + 1], s1); }
            if (c < k) s0 = fma(-q[c], rr[c], s0);
            dst[i] = s0 + s1;
        }
        tm::sync();
    }
    // gs_update plus |dst|^2: the barrier of the team sum is the one that publishes dst (one barrier and one pass over
    // the vector less than gs_update followed by dot)
    __device__ static __noinline__ double gs_update_norm(const double* src, double* dst, int k)
    {
        QP_BIND
        double sq = 0.0;
        for (int i = tid; i < N; i += TEAM) {
            double s0 = src[i], s1 = 0.0;
            const double* q = Q1 + i * LDQ;
            int c = 0;
#pragma unroll 2
            for (; c + 1 < k; c += 2) { s0 = fma(-q[c], rr[c], s0); s1 = fma(-q[c + 1], rr[c + 1], s1); }
            if (c < k) s0 = fma(-q[c], rr[c], s0);
            const double v = s0 + s1;
            dst[i] = v;
            sq = fma(v, v, sq);
        }
        return tm::sum(sq, red);
    }
    __device__ static __forceinline__ void gs_pass(double* v, bool accumulate, int k)
    {
        gs_dots(v, accumulate, k);
        gs_update(v, v, k);
    }

    // rr = RN^-1 d1 as a product with the explicit inverse: thread c sums row c of RI (entries c .. k - 1)
    __device__ static __noinline__ void solve_rn(int k)
    {
        QP_BIND
        if (tid < k) {
            double s0 = 0.0, s1 = 0.0;
            int j = tid;
#pragma unroll 1
            for (; j + 1 < k; j += 2) { s0 = fma(RI[tri(j) + tid], d1[j], s0); s1 = fma(RI[tri(j + 1) + tid], d1[j + 1], s1); }
            if (j < k) s0 = fma(RI[tri(j) + tid], d1[j], s0);
            rr[tid] = s0 + s1;
        }
        tm::sync();
    }

    // Removes active constraint at position l (k active before the call).  RN loses column l and is brought back to
    // triangular form by k - 1 - l Givens rotations of neighbouring rows; the same rotations act on the columns of Q1
    // and of RI (whose row l goes as well).  The sweep over RN runs in warp 0 alone -- lane j owns column j, the
    // rotation of step i comes from lane i by shuffle, so there is no barrier inside the sweep -- and leaves the
    // rotations in gc / gsn; Q1 and RI are then updated by all threads, one row each, in a single pass.
    __device__ static __noinline__ void drop(int l, int k)
    {
        QP_BIND
        static_assert(KMAX <= 32, "one lane per active row");
        const int row = act_row[l];
        double lv = 0.0; int ar = 0, as = 0;
        const bool mv = tid >= l && tid < k - 1;
        if (mv) { lv = lam[tid + 1]; ar = act_row[tid + 1]; as = act_sgn[tid + 1]; }
        if (tid < 32) {
            const int j = tid;
            // columns l + 1 .. k - 1 move one to the left (lane = row); the entry that lands below the diagonal of
            // column c (old (c + 1, c + 1)) stays in a register of lane c
            double sub = 0.0;
#pragma unroll 1
            for (int c = l; c < k - 1; ++c) {
                const double v = j <= c + 1 ? RN[tri(c + 1) + j] : 0.0;
                __syncwarp();
                if (j <= c) RN[tri(c) + j] = v;
                const double t = __shfl_sync(0xffffffffu, v, c + 1);
                if (j == c) sub = t;
            }
            __syncwarp();
            // Givens sweep (lane = column): step i zeroes the sub-diagonal entry of column i with rows i, i + 1
#pragma unroll 1
            for (int i = l; i < k - 1; ++i) {
                double c = 1.0, sn = 0.0;
                if (j == i) {
                    const double a = RN[tri(i) + i], h2 = fma(a, a, sub * sub);
                    const double ih = h2 > 0.0 ? rsqrt(h2) : 0.0;   // (entries are O(1e-3 .. 1e6): no over/underflow of the squares)
                    if (h2 > 0.0) { c = a * ih; sn = sub * ih; }
                    RN[tri(i) + i] = h2 * ih; rdi[i] = ih;
                    gc[i] = c; gsn[i] = sn;
                }
                c = __shfl_sync(0xffffffffu, c, i); sn = __shfl_sync(0xffffffffu, sn, i);
                if (j > i && j < k - 1) {
                    const double ra = RN[tri(j) + i], rb = RN[tri(j) + i + 1];
                    RN[tri(j) + i] = c * ra + sn * rb;
                    RN[tri(j) + i + 1] = -sn * ra + c * rb;
                }
            }
        }
        tm::sync();
        if (mv) { lam[tid] = lv; act_row[tid] = ar; act_sgn[tid] = as; }
        if (tid == 0) { cstate[row] &= 4; st[0] = k - 1; st[1] -= 1; }
        for (int r = tid; r < N; r += TEAM) {                  // Q1 <- Q1 G^T, a row per thread
            double* const q = Q1 + r * LDQ;
            double qa = q[l];
#pragma unroll 1
            for (int i = l; i < k - 1; ++i) {
                const double qb = q[i + 1], c = gc[i], sn = gsn[i];
                q[i] = c * qa + sn * qb;
                qa = -sn * qa + c * qb;
            }
        }
        if (tid < 32) {
            // RI <- (RI without its row l) G^T, leading k - 1 columns.  Lane = old row r (new row rn): columns are
            // rotated in order, column i + 1 is read at step i and written at step i + 1.
            const int r = tid;
            const bool lower = r < l, upper = r > l && r < k;
            const int rn = lower ? r : r - 1;
            double ha = lower ? RI[tri(l) + r] : 0.0;
#pragma unroll 1
            for (int i = l; i < k - 1; ++i) {
                const bool on = lower || (upper && i >= rn);
                const double hb = on ? RI[tri(i + 1) + r] : 0.0;
                __syncwarp();
                if (on) {
                    const double c = gc[i], sn = gsn[i];
                    RI[tri(i) + rn] = c * ha + sn * hb;
                    ha = -sn * ha + c * hb;
                }
            }
        }
        tm::sync();
    }

    // Adds constraint `row` with sign sgn (normal sgn*a, already whitened into w; ww = w.w), current slack sp <= 0.
    // sj0 >= 0: w has three entries, columns sj0 .. sj0 + 2 (force-only row).
    // Returns status; handles partial steps (drops) per Goldfarb-Idnani.
    __device__ static __noinline__ int add_constraint(int row, int sgn, bool is_eq, double sp, double bound_abs, int max_iter,
                                                      double ww, int sj0)
    {
        QP_BIND
        double up = 0.0;
        int k = st[0], nai = st[1], iters = st[2];
#pragma unroll 1
        for (;;) {
            if (iters >= max_iter) { tm::sync(); if (tid == 0) st[2] = iters; tm::sync(); return QPPVM_STATUS_MAX_ITER; }
            double nrm2 = ww;
            if (k > 0) {
                if (sj0 >= 0) gs_dots_sparse(w, sj0, k); else gs_dots(w, false, k);
                nrm2 = gs_update_norm(w, w2, k);
#if QPPVM_SELECTIVE_GS
                if (nrm2 < QPPVM_GS_RATIO * ww)               // Daniel-Gragg-Kaufman-Stewart: a second pass only after cancellation
#endif
                {
                    gs_dots(w2, true, k);                     // CGS2: "twice is enough"
                    nrm2 = gs_update_norm(w2, w2, k);
                }
            } else {
                for (int i = tid; i < N; i += TEAM) w2[i] = w[i];
                tm::sync();
            }
            const bool dependent = !(nrm2 > 1e-22 * ww) || k >= N;
            const bool full = k >= KMAX;
            double t1 = 1e300; int l = -1;
            if (k > 0) solve_rn(k);                            // rr = RN^-1 d1: step of the multipliers, new column of RI
            if (nai > 0) {
                double cand = 1e300; int ci = 0x7fffffff;
                if (tid < k && (act_sgn[tid] & 1) && rr[tid] > 0.0) { cand = lam[tid] / rr[tid]; ci = tid; }
                tm::argmin(cand, ci, red);
                if (ci != 0x7fffffff) { t1 = cand; l = ci; }
            }
            int fail = -1;
            if (dependent && l < 0) {
                // linearly dependent on the working set and nothing can leave: a redundant equality, or an implied
                // inequality whose violation is rounding noise (degenerate vertex of the level-1 feasible set, where
                // the optimality rows pin the level-0 optimum onto its active bounds) -> satisfied within tolerance.
                if (is_eq) fail = (-sp <= 1e-8 * fmax(1.0, bound_abs)) ? QPPVM_STATUS_OK : QPPVM_STATUS_INFEASIBLE;
                else fail = (-sp <= 1e-6 * fmax(1.0, bound_abs)) ? STATUS_IMPLIED : QPPVM_STATUS_INFEASIBLE;
            }
            const double t2 = dependent ? 1e300 : -sp / nrm2;
            const double t = t1 < t2 ? t1 : t2;
            if (fail < 0 && full && t == t2) fail = QPPVM_STATUS_NUMERIC;   // active-set capacity exhausted
            if (fail >= 0) {
                tm::sync();
                if (tid == 0) { st[2] = iters; red[13] = sp; red[14] = bound_abs; red[15] = (double)k; }   // diagnostics
                tm::sync();
                return fail;
            }
            if (nai > 0) { if (tid < k) lam[tid] -= t * rr[tid]; }
            up += t;
            if (!dependent) {
                for (int i = tid; i < N; i += TEAM) u[i] = fma(t, w2[i], u[i]);
                sp = fma(t, nrm2, sp);
            }
            ++iters;
            if (t == t2) {                                    // full step: row becomes active (every thread writes its own entries)
                const double inv = rsqrt(nrm2), nr = nrm2 * inv;
                for (int i = tid; i < N; i += TEAM) Q1[i * LDQ + k] = w2[i] * inv;
                if (tid < k) { RN[tri(k) + tid] = d1[tid]; RI[tri(k) + tid] = -rr[tid] * inv; }
                if (tid == 0) {
                    RN[tri(k) + k] = nr; RI[tri(k) + k] = inv; rdi[k] = inv; lam[k] = up; act_row[k] = row; act_sgn[k] = is_eq ? 2 * sgn : sgn;
                    cstate[row] = (cstate[row] & 4) | ((is_eq || sgn > 0) ? 1 : 3);
                    st[0] = k + 1; st[1] = nai + (is_eq ? 0 : 1); st[2] = iters;
                }
                tm::sync();
                return QPPVM_STATUS_OK;
            }
            tm::sync();                                       // lam and u of every thread are final before the sweep reads them
            drop(l, k);                                       // partial step: blocking constraint leaves
            --k; --nai;
        }
    }

    // Scan the inactive inequality slots [q0, q1) at x; returns the most violated row (-1: none) and publishes its
    // slack (<0), sign and |bound| in red[8..10].  Rows flagged as candidates (cstate bit 2: the working set of the
    // previous level / previous tick) go first: among violated candidates the most violated one, otherwise the most
    // violated row -- any violated row is a valid Goldfarb-Idnani pivot, and the ones that ended up active in a
    // neighbouring problem are rarely dropped again.
    // TAUVAL: the values of the dense slots are kept (tv[q - q0]) for the KKT check and the output recovery.
    __device__ static __noinline__ int scan(int q0, int q1, double* tv)
    {
        QP_BIND
        double worst = 0.0, wkey = 0.0; int widx = 0x7fffffff; int wsgn = 0; double wb = 0.0;
        for (int q = q0 + tid; q < q1; q += TEAM) {
            int r; double val, lo, hi;
            P::eval_slot(rec, ext, ct, q, x, r, val, lo, hi);
            if (P::TAUVAL && tv) tv[q - q0] = val;
            // cstate: 0 inactive, 1 active at lA, 3 active at uA, 2 implied / weakly active.  The side opposite to an
            // active one is still checked: an empty box (lA > uA) must surface as infeasible, not be masked.
            const int cs = cstate[r] & 3;
            const double pri = (cstate[r] & 4) ? 0x1p100 : 1.0;
            if (cs != 2) {
                const double tol = 1e-9 * fmax(1.0, fabs(val));
                const double sl = val - lo, su = hi - val;
                if (cs != 1 && lo > -0.5 * QPPVM_INFTY && sl < -tol && sl * pri < wkey) { worst = sl; wkey = sl * pri; widx = r; wsgn = 1; wb = fabs(lo); }
                if (cs != 3 && hi < 0.5 * QPPVM_INFTY && su < -tol && su * pri < wkey) { worst = su; wkey = su * pri; widx = r; wsgn = -1; wb = fabs(hi); }
            }
        }
        double v = wkey; int idx = widx;
        if (q1 - q0 <= 32) {                                   // all slots live in warp 0: no exchange across warps
            warp_argmin(v, idx);
            if (tid < 32 && widx == idx && wkey == v && idx != 0x7fffffff) { red[8] = worst; red[9] = (double)wsgn; red[10] = wb; red[11] = (double)idx; }
            if (tid == 0 && idx == 0x7fffffff) red[11] = -1.0;
            tm::sync();
            return (int)red[11];                               // (red[8 .. 11] are next written by the next scan: barriers in between)
        }
        tm::argmin(v, idx, red);
        if (idx == 0x7fffffff) return -1;
        if (widx == idx && wkey == v) { red[8] = worst; red[9] = (double)wsgn; red[10] = wb; }   // the owner publishes
        tm::sync();
        return idx;
    }

    __device__ static __noinline__ int add_equalities(int level, int max_iter)
    {
        QP_BIND
        int status = QPPVM_STATUS_OK;
        const int neq = P::n_eq(level);
#pragma unroll 1
        for (int e = 0; e < neq && status == QPPVM_STATUS_OK; ++e) {
            const int row = P::eq_row(level, e);
            double lo, hi;
            P::template build_row<TEAM>(rec, ext, row, eopt, av, lo, hi, tid);
            tm::sync();
            whiten(av, 1.0, w);
            const double s = dot(w, u) - lo;
            const int sgn = s > 0.0 ? -1 : 1;
            if (sgn < 0) for (int i = tid; i < N; i += TEAM) w[i] = -w[i];
            tm::sync();                                        // (also separates the two reductions: Team::sum)
            const double ww = dot(w, w);
            status = add_constraint(row, sgn, true, -fabs(s), fabs(lo), max_iter, ww, -1);
        }
        return status;
    }

    __device__ static __forceinline__ void reset_active_set()
    {
        QP_BIND
        for (int i = tid; i < P::NROWS; i += TEAM) cstate[i] &= 4;   // the candidate flags survive
        if (tid == 0) { st[0] = 0; st[1] = 0; }
        for (int i = tid; i < N; i += TEAM) u[i] = u0[i];
        tm::sync();
    }
    // Start of a level: every row inactive; candidates (bit 2) = the rows active at the end of the previous level of
    // this problem (cstate still holds them) and / or the rows of `wmask` (the working set of the previous tick).
    __device__ static __forceinline__ void init_cstate(int level, const uint32_t* wmask)
    {
        QP_BIND
        for (int i = tid; i < P::NROWS; i += TEAM) {
            bool cand = level > 0 && (cstate[i] & 1);
            if (wmask) cand |= (wmask[i >> 5] >> (i & 31)) & 1u;
            cstate[i] = cand ? 4 : 0;
        }
    }

    // Split shapes: the prepare kernel has orthogonalised the level's equality normals (Q1 columns, RN) and moved the
    // point onto the rows whose right-hand side depends on the record alone (all of it already sits in Q1, RN, rdi
    // and u: solve_level's bulk copies).  Adopt that working set; for level 1
    // finish the six optimality rows (right-hand side = level-0 task value) with their stored factors:
    // slack_e = sum_{c<=e} RN(c,e) (q_c . u) - b_e, u += -(slack_e / RN(e,e)) q_e, the q_c being orthonormal.
    __device__ static __forceinline__ void adopt_equalities(int level)
    {
        QP_BIND
        const int neq = P::n_eq(level);
        if (tid < neq) {
            const int row = P::eq_row(level, tid);
            lam[tid] = 0.0; act_row[tid] = row; act_sgn[tid] = 2; cstate[row] = 1;
        }
        if (tid == 0) { st[0] = neq; st[1] = 0; st[2] += neq; }
        tm::sync();
        if (neq > 6) {
            if (tid < neq) {
                double s0 = 0.0, s1 = 0.0;
                int i = 0;
                for (; i + 1 < N; i += 2) { s0 = fma(Q1[i * LDQ + tid], u[i], s0); s1 = fma(Q1[(i + 1) * LDQ + tid], u[i + 1], s1); }
                if (i < N) s0 = fma(Q1[i * LDQ + tid], u[i], s0);
                d1[tid] = s0 + s1;
            }
            tm::sync();
            if (tid == 0) {
#pragma unroll 1
#pragma unroll 1
                for (int e = 6; e < neq; ++e) {
                    double sl = -eopt[e - 6];
#pragma unroll 1
                    for (int c = 0; c <= e; ++c) sl = fma(RN[tri(e) + c], d1[c], sl);
                    const double dl = -sl * rdi[e];
                    d1[e] += dl; rr[e] = dl;
                }
            }
            tm::sync();
            for (int i = tid; i < N; i += TEAM) {
                double v = u[i];
#pragma unroll 1
                for (int e = 6; e < neq; ++e) v = fma(rr[e], Q1[i * LDQ + e], v);
                u[i] = v;
            }
            tm::sync();
        }
    }

    // Task rows of `level` -> J, u0, jd (xp must hold the proximal centre: zero at the start of a level).
    __device__ static __forceinline__ int load_and_factor(int level, double eps)
    {
        QP_BIND
        const int md = P::template load_tasks<TEAM>(rec, ext, level, Ad, dg, db, tid, *prm_());
        tm::sync();
        // One instantiation serves both levels when level 1 is at most twice as tall (level 0 is padded with zero
        // rows): the kernel is instruction-fetch sensitive, a second copy of the unrolled inversion costs more
        // than a few FMAs on zeros.
        constexpr int MDF0 = (P::MD1 <= 2 * P::MD0) ? P::MD1 : P::MD0;
        if (MDF0 != P::MD0 && level == 0) {
            for (int t = tid; t < (MDF0 - P::MD0) * LDA; t += TEAM) Ad[P::MD0 * LDA + t] = 0.0;
            tm::sync();
        }
        if (level == 0) factor<MDF0>(eps); else factor<P::MD1>(eps);   // Ad dead after this
        return md;
    }

    // One level: returns status.  On return x holds the level solution.
    // The KKT residual of the level is left in red[12].
    __device__ static __noinline__ int solve_level(int level, double eps_reg, int n_reg_steps, int max_iter, double* ydiag,
                                                   const uint32_t* wmask)
    {
        QP_BIND
        const double eps = P::regularised(level) ? eps_reg : 0.0;
        const int steps = eps > 0.0 ? n_reg_steps : 0;
        for (int i = tid; i < N; i += TEAM) xp[i] = 0.0;
        if (tid == 0) st[2] = 0;
        tm::sync();
        int md;
        int status = QPPVM_STATUS_OK;
        if constexpr (P::SPLIT_FACTOR) {
            // everything the prepare kernel produced for this level lands in its shared-memory home through seven
            // bulk copies on one mbarrier (the previous level / problem is done with the slab: barrier above)
            md = level == 0 ? P::MD0 : P::MD1;
            if (tid == 0) {
                const double* wsl = ws_() + level * S::WS_LEVEL;
                constexpr uint32_t B_J = S::SZ_J * 8, B_V = S::VEC * 8, B_Q = QPPVM_WS_COMPACT_Q ? 0 : S::WSZ_Q * 8, B_RN = S::WSZ_RN * 8, B_RDI = S::WSZ_RDI * 8;
                // (compact Q1, staged: the N x n_eq block lands in the tail of the Q1 region, see below)
                const uint32_t B_QS = QPPVM_WS_COMPACT_Q == 2 ? (uint32_t)(((N * P::n_eq(level) + 1) & ~1) * 8) : 0u;
                bulk_expect(mbar_ws_(), (P::J_GLOBAL ? 0 : B_J) + 3 * B_V + B_Q + B_QS + 2 * B_RN + B_RDI);
                if (QPPVM_WS_COMPACT_Q == 2) bulk_copy(Q1 + S::Q_STAGE, wsl + S::WS_Q, B_QS, mbar_ws_());
                if (P::J_GLOBAL) jptr_() = const_cast<double*>(wsl);
                else bulk_copy(Jm, wsl, B_J, mbar_ws_());
                bulk_copy(u0, wsl + S::WS_U0, B_V, mbar_ws_());
                bulk_copy(jd, wsl + S::WS_JD, B_V, mbar_ws_());
                if (!QPPVM_WS_COMPACT_Q) bulk_copy(Q1, wsl + S::WS_Q, B_Q, mbar_ws_());
                bulk_copy(RN, wsl + S::WS_RN, B_RN, mbar_ws_());
                bulk_copy(rdi, wsl + S::WS_RDI, B_RDI, mbar_ws_());
                bulk_copy(u, wsl + S::WS_U, B_V, mbar_ws_());
                bulk_copy(RI, wsl + S::WS_RI, B_RN, mbar_ws_());
            }
#if QPPVM_WS_COMPACT_Q == 1
            {   // the n_eq equality columns of Q1 (compact in the workspace) into their strided home, coalesced loads
                const double* const qws = ws_() + level * S::WS_LEVEL + S::WS_Q;
                static_assert(P::n_eq(1) <= S::NEQ_MAX, "equality columns fit the workspace block");
                if (level == 0) {
                    constexpr int NE = P::n_eq(0);
                    for (int t = tid; t < N * NE; t += TEAM) { const int i = t / NE; Q1[i * LDQ + (t - i * NE)] = qws[t]; }
                } else {
                    constexpr int NE = P::n_eq(1);
                    for (int t = tid; t < N * NE; t += TEAM) { const int i = t / NE; Q1[i * LDQ + (t - i * NE)] = qws[t]; }
                }
            }
#endif
            init_cstate(level, wmask);
            mbar_wait(mbar_ws_(), (uint32_t)st[3]);
            tm::sync();
#if QPPVM_WS_COMPACT_Q == 2
            // compact Q1 (i, c) -> its strided home i LDQ + c.  Source (tail of the Q1 region) and targets overlap: every
            // thread takes its elements into registers, the barrier below separates the reads from the writes
            constexpr int QPT = (N * S::NEQ_MAX + TEAM - 1) / TEAM;
            double qv[QPT];
            {
                const int ne = P::n_eq(level);
#pragma unroll
                for (int m = 0; m < QPT; ++m) { const int t = tid + m * TEAM; qv[m] = t < N * ne ? Q1[S::Q_STAGE + t] : 0.0; }
            }
#endif
            const bool prepared = rdi[S::NEQ_MAX] == 0.0;
            tm::sync();
#if QPPVM_WS_COMPACT_Q == 2
            if (level == 0) {
                constexpr int NE = P::n_eq(0);
#pragma unroll
                for (int m = 0; m < QPT; ++m) { const int t = tid + m * TEAM, i = t / NE; if (t < N * NE) Q1[i * LDQ + (t - i * NE)] = qv[m]; }
            } else {
                constexpr int NE = P::n_eq(1);
#pragma unroll
                for (int m = 0; m < QPT; ++m) { const int t = tid + m * TEAM, i = t / NE; if (t < N * NE) Q1[i * LDQ + (t - i * NE)] = qv[m]; }
            }
#endif
            if (tid == 0) { st[3] ^= 1; st[0] = 0; st[1] = 0; }
            if (prepared && P::n_eq(level) > max_iter) status = QPPVM_STATUS_MAX_ITER;   // as the row-by-row path would
            else if (prepared) adopt_equalities(level);
            else {
                for (int i = tid; i < N; i += TEAM) u[i] = u0[i];
                tm::sync();
                status = add_equalities(level, max_iter);
            }
        } else {
            md = load_and_factor(level, eps);
            init_cstate(level, wmask);
            reset_active_set();
            // ---- equalities first (dyn-feas, then level-0 optimality rows), never dropped
            status = add_equalities(level, max_iter);
        }
        // ---- inequalities + proximal regularisation steps
#pragma unroll 1
        for (int step = 0; status == QPPVM_STATUS_OK; ++step) {
#pragma unroll 1
            for (;;) {
                // Force-only slots first, from u alone (x_f = jd u_f); the triangular product for the full x and the
                // dense (torque-limit) slots only when none of those is violated.  The loop is left with the full x.
                int row = -1;
                if (P::NI_CHEAP > 0) {
                    for (int j = NB + tid; j < N; j += TEAM) x[j] = jd[j] * u[j];
                    tm::sync();
                    row = scan(0, P::NI_CHEAP, nullptr);
                }
                if (row < 0) {
                    unwhiten(u, x);
                    if (P::NI > P::NI_CHEAP) row = scan(P::NI_CHEAP, P::NI, tv_());
                    if (row < 0) break;
                }
                const double sp = red[8], babs = red[10];
                const int sgn = (int)red[9];
                int sj0 = -1;
                double ww, a0, a1, a2;
                if (P::sparse_row(ct, row, sj0, a0, a1, a2)) {    // whitened normal: three entries
                    const double w0 = sgn * jd[sj0] * a0, w1 = sgn * jd[sj0 + 1] * a1, w2v = sgn * jd[sj0 + 2] * a2;
                    for (int i = tid; i < N; i += TEAM) w[i] = i == sj0 ? w0 : (i == sj0 + 1 ? w1 : (i == sj0 + 2 ? w2v : 0.0));
                    tm::sync();
                    ww = fma(w0, w0, fma(w1, w1, w2v * w2v));
                } else {
                    sj0 = -1;
                    double lo, hi;
                    P::template build_row<TEAM>(rec, ext, row, eopt, av, lo, hi, tid);
                    tm::sync();
                    whiten(av, (double)sgn, w);
                    ww = dot(w, w);
                }
                status = add_constraint(row, sgn, false, sp, babs, max_iter, ww, sj0);
                if (status == STATUS_IMPLIED) {               // not added; excluded from further scans
                    if (tid == 0) cstate[row] = (cstate[row] & 4) | 2;
                    tm::sync();
                    status = QPPVM_STATUS_OK;
                }
                if (status != QPPVM_STATUS_OK) break;
            }
            if (status != QPPVM_STATUS_OK || step >= steps) break;
            // qpOASES solveRegularisedQP(): g <- g_orig - eps x_prev, i.e. u0 += delta, delta = eps J^T (x - xp_old);
            // same active set: u += (I - Q1 Q1^T) delta, lam -= RN^-1 Q1^T delta.
            for (int i = tid; i < N; i += TEAM) { av[i] = eps * (x[i] - xp[i]); xp[i] = x[i]; }
            tm::sync();
            whiten(av, 1.0, w);
            for (int i = tid; i < N; i += TEAM) { u0[i] += w[i]; w2[i] = w[i]; }
            tm::sync();
            const int k = st[0];
            if (k > 0) {
                gs_pass(w2, false, k);
                gs_pass(w2, true, k);
                solve_rn(k);
                if (tid < k) lam[tid] -= rr[tid];
                // (equality multipliers are re-derived at output time from u - u0); inequality ones must stay >= 0
                const bool neg = tid < k && (act_sgn[tid] & 1) && lam[tid] < 0.0;
                if (tm::any(neg)) {
                    // rare: active set changes under the proximal shift -> cold restart of this step from u0
                    reset_active_set();
                    status = add_equalities(level, max_iter);
                    continue;
                }
            }
            for (int i = tid; i < N; i += TEAM) u[i] += w2[i];
            tm::sync();
        }
        if (status == QPPVM_STATUS_OK) {                       // non-finite data must not reach the command
            bool bad = false;
            for (int i = tid; i < N; i += TEAM) bad |= !isfinite(x[i]);
            if (tm::any(bad)) status = QPPVM_STATUS_NUMERIC;
        }
        if (status != QPPVM_STATUS_OK) return status;
        if constexpr (P::SPLIT_FACTOR) {
            // The KKT certificate of these shapes is computed by qp_certify_kernel from the record and the block
            // exported here (the check is regular code that only pollutes this kernel's instruction cache).
            (void)md;
            export_certificate(level, ydiag);
            if (tid == 0) red[12] = (double)__int_as_float(0x7f800000);
        } else {
            const double kv = kkt(level, md, eps, ydiag);
            if (tid == 0) red[12] = kv;
        }
        tm::sync();
        return status;
    }

    // Signed multipliers y (qpOASES convention: > 0 active at lA, < 0 at uA) of the level's final working set from
    // u - u0 = sum lam_c w_c (RN lam = Q1^T (u - u0)), and the block qp_certify_kernel reads.  Rows whose multiplier
    // is exactly zero are weakly active: not reported in the active mask.
    __device__ static __noinline__ void export_certificate(int level, double* ydiag)
    {
        QP_BIND
        const int k = st[0];
        double* const cb = ws_() + level * S::WS_LEVEL;
        for (int i = tid; i < N; i += TEAM) { w2[i] = u[i] - u0[i]; cb[S::C_X + i] = x[i]; cb[S::C_XP + i] = xp[i]; }
        if (tid < QPPVM_M0) cb[S::C_EOPT + tid] = eopt[tid];
        tm::sync();
        if (k > 0) { gs_dots(w2, false, k); solve_rn(k); }
        int* const ci = reinterpret_cast<int*>(cb + S::C_INT);
        if (tid < k) {
            const int row = act_row[tid], sg = act_sgn[tid];
            const double y = (sg > 0 ? 1.0 : -1.0) * rr[tid];
            cb[S::C_Y + tid] = y;
            ci[1 + tid] = row | (sg << 8);                   // sign: +-1 inequality, +-2 equality
            if ((sg & 1) && y == 0.0) cstate[row] = (cstate[row] & 4) | 2;
            if (ydiag) ydiag[row] = y;
        }
        if (tid == 0) ci[0] = k;
        tm::sync();
    }

    // bits [32 wd, 32 wd + 32) of the active-row mask (rows whose multiplier is non-zero)
    __device__ static __noinline__ uint32_t active_word(int wd)
    {
        uint32_t mask = 0;
#pragma unroll 1
        for (int b = 0; b < 32; ++b) {
            const int r = wd * 32 + b;
            if (r < P::NROWS && (cstate_()[r] & 1)) mask |= 1u << b;
        }
        return mask;
    }

    // KKT certificate of the solved (regularised, proximal-shifted) level problem, SURVEY.md 8(c).
    __device__ static __noinline__ double kkt(int level, int md, double eps, double* ydiag)
    {
        QP_BIND
        const int k = st[0];
        // Signed multipliers y (qpOASES convention: > 0 active at lA, < 0 at uA) from u - u0 = sum lam_c w_c:
        // RN lam = Q1^T (u - u0); recomputed here so that equality multipliers are exact after all updates.
        for (int i = tid; i < N; i += TEAM) w2[i] = u[i] - u0[i];
        tm::sync();
        if (tid < k) {
            double s = 0.0;
#pragma unroll 2
            for (int i = 0; i < N; ++i) s = fma(Q1[i * LDQ + tid], w2[i], s);
            d1[tid] = s;
        }
        tm::sync();
        solve_rn(k);                                          // rr = multipliers of the signed normals
        // grad = D (x - db) + Ad^T (Ad x - b) + eps (x - xp) - sum_c y_c a_c ; constraint part first (w)
        for (int i = tid; i < N; i += TEAM) w[i] = 0.0;
        tm::sync();
        double rprim = 0.0, rcomp = 0.0, cxmax = 0.0, ymax = 0.0;   // per-thread partial maxima, merged at the end
        int c0 = 0;
        if constexpr (P::HAS_EQ_COEF) {
            // The equality rows sit at the head of the working set in row order (unless a dependent one was skipped):
            // one sweep with every thread fetching its column of all of them (loads in flight together), the
            // a_e . x through one transposed reduction, instead of a row build + reduction + barriers per row.
            const int neq = P::n_eq(level);
            bool fast = k >= neq;
            for (int e = 0; e < neq && fast; ++e) fast = act_row[e] == P::eq_row(level, e) && !(act_sgn[e] & 1);
            if (fast) {
                static_assert(S::NEQ_MAX <= 16 && TEAM == 64, "one transposed reduction per warp");
                double pe[16], wj = 0.0;
                const double xj = tid < N ? x[tid] : 0.0;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    double a = 0.0;
                    if (e < neq && tid < N) {
                        double sg;
                        const int off = P::eq_coef(level, e, tid, sg);
                        if (off >= 0) a = sg * ext[off];
                    }
                    const double y = e < neq ? (act_sgn[e] > 0 ? 1.0 : -1.0) * rr[e] : 0.0;
                    wj = fma(-y, a, wj);
                    pe[e] = a * xj;
                }
                const double v = warp_sum_transposed<16>(pe, tid & 31);
                if (!(tid & 1)) av[(tid >> 5) * 16 + ((tid & 31) >> 1)] = v;
                if (tid < N) w[tid] = wj;
                tm::sync();
                if (tid < neq) {
                    const double val = av[tid] + av[16 + tid];
                    const double y = (act_sgn[tid] > 0 ? 1.0 : -1.0) * rr[tid];
                    rprim = fabs(val - P::eq_bound(rec, tid, eopt)); cxmax = fabs(val); ymax = fabs(y);
                    if (ydiag) ydiag[act_row[tid]] = y;
                }
                tm::sync();
                c0 = neq;
            }
        }
#pragma unroll 1
        for (int c = c0; c < k; ++c) {
            const int row = act_row[c];
            const int sg = act_sgn[c];
            double lo, hi;
            P::template build_row<TEAM>(rec, ext, row, eopt, av, lo, hi, tid);
            tm::sync();
            const bool iseq = !(sg & 1);                       // sign = +-1, equalities +-2
            const double y = (sg > 0 ? 1.0 : -1.0) * rr[c];
            const double val = dot(av, x);
            for (int i = tid; i < N; i += TEAM) w[i] = fma(-y, av[i], w[i]);
            cxmax = fmax(cxmax, fabs(val)); ymax = fmax(ymax, fabs(y));
            if (iseq) rprim = fmax(rprim, fabs(val - lo));
            else {
                rprim = fmax(rprim, fmax(0.0, fmax(lo - val, val - hi)));
                rcomp = fmax(rcomp, y > 0.0 ? y * fabs(val - lo) : -y * fabs(hi - val));
                if (sg * y < 0.0) rcomp = fmax(rcomp, fabs(y));        // wrong-signed multiplier
                if (y == 0.0 && tid == 0) cstate[row] = (cstate[row] & 4) | 2;   // weakly active: not reported in the mask
            }
            if (ydiag && tid == 0) ydiag[row] = y;
            tm::sync();
        }
        // inactive inequalities: primal violation only
        {
            double viol = 0.0, cm = 0.0;
            for (int q = tid; q < P::NI; q += TEAM) {
                int r; double val, lo, hi;
                if (P::TAUVAL && q >= P::NI_CHEAP) { P::dense_slot_bounds(rec, q, r, lo, hi); val = tv_()[q - P::NI_CHEAP]; }
                else P::eval_slot(rec, ext, ct, q, x, r, val, lo, hi);
                cm = fmax(cm, fabs(val));
                if (!(cstate[r] & 1)) viol = fmax(viol, fmax(lo - val, val - hi));
            }
            rprim = fmax(rprim, viol); cxmax = fmax(cxmax, cm);
        }
        // task part (reload the dense task rows over the dead Q1 region); (A x)_r -> w2, b_r -> av
        tm::sync();
        P::template load_tasks<TEAM>(rec, ext, level, Ad, dg, db, tid, *prm_());
        tm::sync();
        for (int r = tid; r < md; r += TEAM) {
            double s0 = 0.0, s1 = 0.0;
            int j = 0;
#pragma unroll 2
            for (; j + 1 < NB; j += 2) { s0 = fma(Ad[r * LDA + j], x[j], s0); s1 = fma(Ad[r * LDA + j + 1], x[j + 1], s1); }
            if (j < NB) s0 = fma(Ad[r * LDA + j], x[j], s0);
            w2[r] = s0 + s1;
            av[r] = Ad[r * LDA + NB];
        }
        tm::sync();
        double rs = 0.0, gmax = 0.0, hxmax = 0.0, xmax = 0.0;
        for (int j = tid; j < N; j += TEAM) {
            double hx = (dg[j] + eps) * x[j], g = -dg[j] * db[j] - eps * xp[j];
            if (j < NB)
#pragma unroll 2
                for (int r = 0; r < md; ++r) { hx = fma(Ad[r * LDA + j], w2[r], hx); g = fma(-Ad[r * LDA + j], av[r], g); }
            const double stn = hx + g + w[j];
            rs = fmax(rs, fabs(stn)); gmax = fmax(gmax, fabs(g)); hxmax = fmax(hxmax, fabs(hx)); xmax = fmax(xmax, fabs(x[j]));
        }
        {   // the eight maxima meet through one exchange (red: WARPS x 8)
            static_assert(tm::WARPS * 8 <= 16, "red holds the exchange");
            double m[8] = {rs, gmax, hxmax, xmax, rprim, rcomp, cxmax, ymax};
#pragma unroll
            for (int q = 0; q < 8; ++q) m[q] = warp_max(m[q]);
            tm::sync();
            if ((tid & 31) == 0)
#pragma unroll
                for (int q = 0; q < 8; ++q) red[(tid >> 5) * 8 + q] = m[q];
            tm::sync();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                double v = red[q];
#pragma unroll
                for (int wv = 1; wv < tm::WARPS; ++wv) v = fmax(v, red[wv * 8 + q]);
                m[q] = v;
            }
            tm::sync();
            rs = m[0]; gmax = m[1]; hxmax = m[2]; xmax = m[3]; rprim = m[4]; rcomp = m[5]; cxmax = m[6]; ymax = m[7];
        }
        rs /= fmax(1.0, fmax(gmax, hxmax));
        rprim /= fmax(1.0, fmax(xmax, cxmax));
        rcomp /= fmax(1.0, ymax) * fmax(1.0, cxmax);
        return fmax(rs, fmax(rprim, rcomp));
    }
#undef QP_BIND
};

// ------------------------------------------------------------------------------------------
// Kernel: persistent CTAs (one team each) pull problem indices from a global counter.
// ------------------------------------------------------------------------------------------
template <class P, int TEAM>
__global__ void __launch_bounds__(TEAM, Slab<P>::CTAS)
qp_solve_kernel(const double* __restrict__ recs, unsigned char* __restrict__ out, double* __restrict__ diag,
                long long batch, Params prm, unsigned long long* __restrict__ counter, double* __restrict__ ws,
                uint32_t* __restrict__ warm, Tick tk)
{
    // warm (optional, in/out): 8 words per problem, the active rows of level 0 | level 1 at the end of the previous
    // solve of this problem (previous control tick; zeros = cold).  They are tried first (scan()), which is what
    // the reference gets from keeping one QPOases_sot alive across ticks (ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:64,
    // ref:src/QPPVMPlugin.cpp:246: qpOASES hot start).
    using SV = Solver<P, TEAM>;
    constexpr int N = P::N;
    const int tid = threadIdx.x;
    constexpr int OUT_BYTES = 8 * (N + P::NA) + 32;
    constexpr int DIAG = N + 2 * P::NROWS + QPPVM_M0;
    __shared__ unsigned long long s_idx;
    if (tid == 0) { mbar_init(SV::mbar_(), 1); mbar_init(SV::mbar_ws_(), 1); SV::state_()[3] = 0; *SV::prm_() = prm; }   // state[3]: phase of the workspace copies
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    unsigned long long next_static = blockIdx.x;               // counter == nullptr: static round-robin schedule
    uint32_t tick = tk.seq0;
#pragma unroll 1
    for (;;) {
        if (tid == 0) {
            unsigned long long i;
            if (tk.host) {                                     // resident: problem 0 again once the prepare server is done
                i = tick_wait_dev(tk, 0, 2, tick) ? 0ull : (unsigned long long)batch;
                asm volatile("fence.proxy.async;" ::: "memory");   // its generic-proxy writes vs. the bulk copies below
                tick_stamp(tk, 3);
            }
            else if (counter) i = atomicAdd(counter, 1ull);
            else { i = next_static; next_static += gridDim.x; }
            s_idx = i;
            if ((long long)i < batch) {
                const double* gr = recs + i * (size_t)P::REC;
                *reinterpret_cast<const double**>(reinterpret_cast<double*>(g_smem) + Slab<P>::O_J - 2) = gr;
                if (P::SPLIT_FACTOR) {
                    SV::ws_() = ws + i * (size_t)Slab<P>::WS;
                    // level 1's share of the workspace is needed ~100 us from now: have it in L2 by then
                    l2_prefetch(ws + i * (size_t)Slab<P>::WS + Slab<P>::WS_LEVEL, (uint32_t)(Slab<P>::WS_LEVEL * 8));
                    // the record was pushed out of L2 by the workspace writes of the prepare kernel; the row / task
                    // reads of the inequality scan, the KKT check and the torque recovery all come from it
                    if (!Slab<P>::STAGE) l2_prefetch(gr, (uint32_t)(P::REC * sizeof(double)));
                }
                if (Slab<P>::STAGE)                            // one TMA bulk copy stages the tail of the record
                    bulk_load(SV::rec_(), gr + P::STAGE_FROM, (uint32_t)((P::REC - P::STAGE_FROM) * sizeof(double)), SV::mbar_());
            }
        }
        __syncthreads();
        const unsigned long long idx = s_idx;
        if ((long long)idx >= batch) break;
        double* xo = reinterpret_cast<double*>(out + idx * (size_t)OUT_BYTES);
        double* dg = diag ? diag + idx * (size_t)DIAG : nullptr;
        uint32_t* const wm = warm ? warm + idx * 8 : nullptr;
        if (dg) for (int i = tid; i < DIAG; i += TEAM) dg[i] = 0.0;
        if (Slab<P>::STAGE) {
            // linear contact-Jacobian rows next to the staged tail (plain coalesced loads, overlapping the TMA copy)
            const double* gr = recs + idx * (size_t)P::REC;
            for (int t = tid; t < P::WD * P::NC * P::NV; t += TEAM) {
                const int rowl = t / P::NV, col = t - rowl * P::NV;
                SV::rec_()[P::S_JCL + t] = gr[P::NV * 6 + ((rowl / P::WD) * 6 + (rowl % P::WD)) * P::NV + col];
            }
            mbar_wait(SV::mbar_(), phase); phase ^= 1;
            __syncthreads();
        }
        P::template fill_slot_table<TEAM>(SV::rec_(), SV::ct_(), tid);    // (visible after the first barrier of solve_level)
        float kkt0 = __int_as_float(0x7f800000), kkt1 = kkt0;
        int it0 = 0, it1 = 0;
        int status = P::template prepare<TEAM>(SV::rec_(), P::EXT_IS_GLOBAL ? SV::grec_() : SV::ext_(), tid) ? QPPVM_STATUS_OK : QPPVM_STATUS_NUMERIC;
        if (status == QPPVM_STATUS_OK)
            status = SV::solve_level(0, prm.eps_reg, prm.n_reg_steps, prm.max_iter, dg ? dg + N : nullptr, wm);
        it0 = SV::state_()[2];
        if (wm && status == QPPVM_STATUS_OK && tid < 4) wm[tid] = SV::active_word(tid);
        if (status == QPPVM_STATUS_OK) kkt0 = (float)SV::red_()[12];
        if (status == QPPVM_STATUS_OK) {
            P::template task0_value<TEAM>(SV::rec_(), P::EXT_IS_GLOBAL ? SV::grec_() : SV::ext_(), SV::x_(), SV::eopt_(), tid);
            Team<TEAM>::sync();
            if (dg) {
                for (int i = tid; i < N; i += TEAM) dg[i] = SV::x_()[i];
                if (tid < QPPVM_M0) dg[N + 2 * P::NROWS + tid] = SV::eopt_()[tid];
            }
            Team<TEAM>::sync();
            status = SV::solve_level(1, prm.eps_reg, prm.n_reg_steps, prm.max_iter, dg ? dg + N + P::NROWS : nullptr, wm ? wm + 4 : nullptr);
            it1 = SV::state_()[2];
            if (status == QPPVM_STATUS_OK) kkt1 = (float)SV::red_()[12];
        }
        const bool ok = status == QPPVM_STATUS_OK;
        if (dg && !ok && tid < 3) dg[N + 2 * P::NROWS + 3 + tid] = SV::red_()[13 + tid];   // slack, |bound|, k at failure
        for (int i = tid; i < N; i += TEAM) xo[i] = ok ? SV::x_()[i] : 0.0;
        P::template recover<TEAM>(SV::rec_(), P::EXT_IS_GLOBAL ? SV::grec_() : SV::ext_(), SV::x_(), SV::tv_(), xo + N, ok, tid);
        // trailer: status, iters, 128-bit active mask of level 1, kkt[2]
        uint32_t* tr = reinterpret_cast<uint32_t*>(xo + N + P::NA);
        if (tid < 4) {
            const uint32_t mask = ok ? SV::active_word(tid) : 0u;
            tr[2 + tid] = mask;
            if (wm && ok) wm[4 + tid] = mask;
        }
        if (tid == 0) {
            tr[0] = (uint32_t)status; tr[1] = (uint32_t)((it0 & 0xffff) | (it1 << 16));
            tr[6] = __float_as_uint(kkt0); tr[7] = __float_as_uint(kkt1);
        }
        __syncthreads();                                       // slab (incl. s_idx, rec) is reused by the next problem
        if (tk.host) {
            __threadfence();
            if (tid == 0) { tick_stamp(tk, 4); st_release_gpu(tk.dev + 1, tick); }
        }
    }
    if (tk.host && tid == 0) st_release_gpu(tk.dev + 3, 1u);   // the certify server follows
}

// ------------------------------------------------------------------------------------------
// Factor kernel (shapes with P::SPLIT_FACTOR).  The factorisation has no data-dependent control flow, so several
// (problem, level) pairs run side by side in one CTA: NB + 1 lanes per pair (one per task column + the rhs column),
// FactorShape::FPC pairs per CTA -- 7 x 36 = 252 of 256 lanes for the 29-DoF shape, where a 64-thread team leaves
// 28 lanes (and almost all of its second warp) idle.  Grid-stride over groups of FPC pairs.  Writes J | u0 | jd of
// each level into the workspace the solve kernel reads.
// ------------------------------------------------------------------------------------------
// dmine (lane c) = q_c . v for c < e, the e dot products of one Gram-Schmidt pass (rows i0, i1 of this lane)
template <int NV, int N>
__device__ __forceinline__ double gs_dots(const double* Wq, int e, int i0, int i1, bool has1, double va, double vb, int l)
{
    double p[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        const int cc = c < e ? c : 0;
        const double qa = Wq[cc * N + i0], qb = has1 ? Wq[cc * N + i1] : 0.0;
        p[c] = fma(qa, va, qb * vb);
    }
    const double v = warp_sum_transposed<NV>(p, l);
    return __shfl_sync(0xffffffffu, v, (l * (32 / NV)) & 31);      // lane c <- the lanes holding value c
}

template <class P>
struct FactorShape {
    static constexpr int N = P::N, NB = P::NB, GS = NB + 1, MD = P::MD1 > P::MD0 ? P::MD1 : P::MD0;
    static constexpr int THREADS = GS <= 36 ? 256 : 128;        // the 51-variable shapes need 22 KB per pair: 3 pairs per CTA
    static constexpr int FPC = THREADS / GS;
    static constexpr int VEC = Slab<P>::VEC, SZ_J = Slab<P>::SZ_J, LDA = NB + 1;
    // per-pair block (doubles): J | Ad | dg | db | u0 | jd | broadcast | normals -> Q | RN | 1/diag.
    // After the factorisation Ad|dg|db hold the equality rows [e][i] and the broadcast slots their right-hand sides.
    static constexpr int NEQ = Slab<P>::NEQ_MAX;
    // (Ad | dg | db are reused for the NEQ x N equality rows: wide shapes with few task rows pad Ad)
    static constexpr int SZ_AD_RAW = MD * LDA > NEQ * N - 2 * VEC ? MD * LDA : NEQ * N - 2 * VEC;
    static constexpr int O_AD = SZ_J, O_DG = O_AD + SZ_AD_RAW + (SZ_AD_RAW & 1), O_DB = O_DG + VEC, O_U0 = O_DB + VEC;
#if QPPVM_PREP_INPLACE
    // the whitened normals overwrite the equality rows in place (every lane keeps its column of all rows in registers
    // across one barrier): no separate NEQ x N block
    static constexpr int O_JD = O_U0 + VEC, O_BC = O_JD + VEC, O_WQ = O_AD;
    static constexpr int O_RN = O_BC + 2 * (MD + 4), O_RDI = O_RN + NEQ * NEQ, BLOCK = O_RDI + NEQ;
#else
    static constexpr int O_JD = O_U0 + VEC, O_BC = O_JD + VEC, O_WQ = O_BC + 2 * (MD + 4);
    static constexpr int O_RN = O_WQ + NEQ * N + ((NEQ * N) & 1), O_RDI = O_RN + NEQ * NEQ, BLOCK = O_RDI + NEQ;
#endif
    static_assert(O_U0 - O_AD >= NEQ * N, "equality rows fit over Ad | dg | db");
    static_assert(2 * (MD + 4) >= NEQ, "right-hand sides fit in the broadcast slots");
    static_assert(FPC <= THREADS / 32, "one warp per pair in the orthogonalisation phase");
    static constexpr int BYTES = FPC * BLOCK * 8;
    // resident CTAs per SM the shared memory allows (the register budget follows through __launch_bounds__), at most 4
    static constexpr int CTAS_RAW = 233472 / (BYTES + 1024);
    static constexpr int CTAS = QPPVM_PREP_INPLACE ? (CTAS_RAW > 4 ? 4 : (CTAS_RAW < 1 ? 1 : CTAS_RAW)) : 1;
};

template <class P>
__global__ void __launch_bounds__(FactorShape<P>::THREADS, FactorShape<P>::CTAS)
qp_factor_kernel(const double* __restrict__ recs, double* __restrict__ ws, long long batch, Params prm,
                 unsigned long long* __restrict__ counter, Tick tk)
{
    using F = FactorShape<P>;
    using S = Slab<P>;
    constexpr int N = P::N;
    const int t = threadIdx.x;
    // the work counter of the solve kernel that follows on this stream (the previous solve is complete: stream order)
    if (counter && blockIdx.x == 0 && t == 0) *counter = 0ull;
    const int f = t / F::GS, lane = t - f * F::GS;
    double* const blk = reinterpret_cast<double*>(g_smem) + (f < F::FPC ? f : 0) * F::BLOCK;
    double* const Jm = blk; double* const Ad = blk + F::O_AD; double* const dg = blk + F::O_DG;
    double* const db = blk + F::O_DB; double* const u0 = blk + F::O_U0; double* const jd = blk + F::O_JD;
    double* const bc = blk + F::O_BC;
    __shared__ uint32_t s_cmd;
    uint32_t tick = tk.seq0;
#pragma unroll 1
    for (long long base = (long long)blockIdx.x * F::FPC; base < 2 * batch; base += (long long)gridDim.x * F::FPC) {
        if (tk.host) {
            // resident: wait for the host to post the next tick, pull the record across PCIe into device memory
            if (t == 0) {
                const unsigned long long t0 = globaltimer_ns();
                uint32_t cmd;
                for (;;) {
                    cmd = ld_acquire_sys(tk.host);
                    if (cmd != tick) break;
                    if (ld_acquire_sys(tk.host + 2) != 0u || globaltimer_ns() - t0 > 1000ull * tk.idle_us) { cmd = tick; break; }
                }
                s_cmd = cmd;
            }
            __syncthreads();
            const uint32_t cmd = s_cmd;
            __syncthreads();
            if (cmd == tick) {                                 // quit request or idle for too long: take the chain down
                if (t == 0) st_release_gpu(tk.dev + 2, 1u);
                break;
            }
            tick = cmd;
            if (t == 0) tick_stamp(tk, 0);
            for (int i = t; i < P::REC / 2; i += F::THREADS)
                reinterpret_cast<double2*>(tk.dev_rec)[i] = __ldcv(reinterpret_cast<const double2*>(tk.host_rec) + i);
            __syncthreads();
            if (t == 0) tick_stamp(tk, 1);
            base = 0;
        }
        // pairs are ordered level-major ([0, batch): level 0, [batch, 2 batch): level 1) so that the pairs sharing a
        // CTA pass have the same number of equality rows (the orthogonalisation phase is 4x longer for level 1)
        const long long pair = base + f;
        const bool live = f < F::FPC && pair < 2 * batch;
        const int level = pair >= batch ? 1 : 0;
        const long long idx = pair - (level ? batch : 0);
        if (live) {
            const double* gr = recs + idx * (size_t)P::REC;
            // (policy functions address the staged tail as rec[OFF - SB]: hand them the global record shifted by SB)
            const int md = P::template load_tasks<F::GS>(gr + P::SB, gr, level, Ad, dg, db, lane, prm);
            for (int e = md * F::LDA + lane; e < F::MD * F::LDA; e += F::GS) Ad[e] = 0.0;   // pad to the common height
        }
        __syncthreads();
        const double eps = P::regularised(level) ? prm.eps_reg : 0.0;
        // Householder QR with as many dense rows as the pass needs: a pass whose pairs are all level 0 (pairs are
        // level-major) runs the 6-row version instead of the padded one (the stride of Ad stays LDA either way)
        const bool all_l0 = base + F::FPC <= batch;            // CTA-uniform
        if (P::MD0 < F::MD && all_l0) factor_qr<P::MD0, N, P::NB, F::GS>(live ? lane : -1, Jm, Ad, dg, db, u0, jd, bc, eps);
        else factor_qr<F::MD, N, P::NB, F::GS>(live ? lane : -1, Jm, Ad, dg, db, u0, jd, bc, eps);
        factor_backsub<P::NB>(live ? lane : -1, Jm, jd);
        double* const wsl = ws + idx * (size_t)S::WS + level * S::WS_LEVEL;
        double* const Aeq = blk + F::O_AD; double* const lo_eq = bc;
        double* const WQ = blk + F::O_WQ; double* const RNb = blk + F::O_RN; double* const rdib = blk + F::O_RDI;
        const int neq = live ? P::n_eq(level) : 0;
        if (live) {
            for (int i = lane; i < S::SZ_J; i += F::GS) wsl[i] = Jm[i];
            for (int i = lane; i < N; i += F::GS) { wsl[S::WS_U0 + i] = u0[i]; wsl[S::WS_JD + i] = jd[i]; }
            // equality rows of the level: each lane fetches its column of all rows before storing any of them, so
            // the global loads of a pass are in flight together
            const double* gr = recs + idx * (size_t)P::REC;
            for (int jc = lane; jc < N; jc += F::GS) {
                double v[F::NEQ];
#pragma unroll
                for (int e = 0; e < F::NEQ; ++e) {
                    double sg;
                    const int off = e < neq ? P::eq_coef(level, e, jc, sg) : -1;
                    v[e] = off >= 0 ? sg * gr[off] : 0.0;
                }
#pragma unroll
                for (int e = 0; e < F::NEQ; ++e) Aeq[e * N + jc] = v[e];
            }
            if (lane < neq) lo_eq[lane] = P::eq_rhs(gr, level, lane);
        }
        __syncthreads();
        // whitened normals w_e = J^T a_e, all rows at once (lane j: column j of J against every row)
        {
            double acc[F::NEQ];
#pragma unroll
            for (int e = 0; e < F::NEQ; ++e) acc[e] = 0.0;
            if (live && lane < P::NB) {
                const double* col = Jm + lane * (lane + 1) / 2;
#pragma unroll 1
                for (int i = 0; i <= lane; ++i) {
                    const double jv = col[i];
#pragma unroll
                    for (int e = 0; e < F::NEQ; ++e) acc[e] = fma(jv, Aeq[e * N + i], acc[e]);
                }
            }
            if (F::O_WQ == F::O_AD) __syncthreads();           // in place: every lane has read the rows it needs
            if (live) {
                if (lane < P::NB) {
#pragma unroll
                    for (int e = 0; e < F::NEQ; ++e) WQ[e * N + lane] = acc[e];
                }
                for (int jj = P::NB + lane; jj < N; jj += F::GS) {
                    const double dj = jd[jj];
#pragma unroll
                    for (int e = 0; e < F::NEQ; ++e) WQ[e * N + jj] = dj * Aeq[e * N + jj];
                }
            }
        }
        __syncthreads();
        // Orthogonalisation (classical Gram-Schmidt, two passes) and the moves onto the rows whose right-hand side
        // is known: one warp per pair, rows i = l and l + 32 per lane, Q overwrites the normals column by column.
        // The same arithmetic as Solver::add_constraint for an equality; a dependent row (or non-finite data)
        // raises the flag and the solve kernel takes the row-by-row path for that problem.
        {
            const int wq = t >> 5, l = t & 31;
            const long long wpair = base + wq;
            if (wq < F::FPC && wpair < 2 * batch) {
                double* const bw = reinterpret_cast<double*>(g_smem) + wq * F::BLOCK;
                double* const Wq = bw + F::O_WQ; double* const Rq = bw + F::O_RN; double* const rdq = bw + F::O_RDI;
                const double* const u0q = bw + F::O_U0; const double* const loq = bw + F::O_BC;
                const int wlevel = wpair >= batch ? 1 : 0, wneq = P::n_eq(wlevel);
                const int nknown = wlevel == 0 ? wneq : 6;       // level 1: rows >= 6 wait for the level-0 task value
                double* const wso = ws + (wpair - (wlevel ? batch : 0)) * (size_t)S::WS + wlevel * S::WS_LEVEL;
                const int i0 = l, i1 = l + 32;
                const bool has1 = i1 < N;
                double ua = u0q[i0], ub = has1 ? u0q[i1] : 0.0;
                double flag = prm.rowwise ? 1.0 : 0.0;
#pragma unroll 1
                for (int e = 0; e < (prm.rowwise ? 0 : wneq); ++e) {
                    const double wa = Wq[e * N + i0], wb = has1 ? Wq[e * N + i1] : 0.0;
                    const double ww = warp_sum(fma(wa, wa, wb * wb));
                    double va = wa, vb = wb, dacc = 0.0;
#pragma unroll 1
                    for (int pass = 0; pass < 2; ++pass) {
                        double dmine = 0.0;
                        if (e > 8) dmine = gs_dots<16, N>(Wq, e, i0, i1, has1, va, vb, l);
                        else if (e > 4) dmine = gs_dots<8, N>(Wq, e, i0, i1, has1, va, vb, l);
                        else if (e > 0) dmine = gs_dots<4, N>(Wq, e, i0, i1, has1, va, vb, l);
                        double na = va, nb = vb, na2 = 0.0, nb2 = 0.0;
#pragma unroll 1
                        for (int c = 0; c + 1 < e; c += 2) {
                            const double dc0 = __shfl_sync(0xffffffffu, dmine, c), dc1 = __shfl_sync(0xffffffffu, dmine, c + 1);
                            na = fma(-dc0, Wq[c * N + i0], na); na2 = fma(-dc1, Wq[(c + 1) * N + i0], na2);
                            if (has1) { nb = fma(-dc0, Wq[c * N + i1], nb); nb2 = fma(-dc1, Wq[(c + 1) * N + i1], nb2); }
                        }
                        if (e & 1) {
                            const double dc = __shfl_sync(0xffffffffu, dmine, e - 1);
                            na = fma(-dc, Wq[(e - 1) * N + i0], na);
                            if (has1) nb = fma(-dc, Wq[(e - 1) * N + i1], nb);
                        }
                        va = na + na2; vb = nb + nb2; dacc += dmine;
                    }
                    const double nrm2 = warp_sum(fma(va, va, vb * vb));
                    if (!(nrm2 > 1e-22 * ww)) { flag = 1.0; break; }
                    const double nr = sqrt(nrm2), inv = 1.0 / nr;
                    __syncwarp();
                    Wq[e * N + i0] = va * inv;
                    if (has1) Wq[e * N + i1] = vb * inv;
                    if (l < e) Rq[e * F::NEQ + l] = dacc;
                    if (l == e) { Rq[e * F::NEQ + e] = nr; rdq[e] = inv; }
                    if (e < nknown) {
                        const double sl = warp_sum(fma(wa, ua, wb * ub)) - loq[e];
                        const double tstep = -sl / nrm2;
                        ua = fma(tstep, va, ua); ub = fma(tstep, vb, ub);
                    }
                    __syncwarp();
                }
                for (int t2 = l; t2 < wneq * N; t2 += 32) {      // Q1 row-major: (i, c) -> i n_eq + c (compact) or its shared-memory home i LDQ + c
                    const int i = t2 / wneq, c = t2 - i * wneq;
                    wso[S::WS_Q + (QPPVM_WS_COMPACT_Q ? t2 : i * S::LDQ + c)] = Wq[c * N + i];
                }
                for (int t2 = l; t2 < wneq * F::NEQ; t2 += 32) {
                    const int c = t2 / F::NEQ, r = t2 - c * F::NEQ;
                    if (r <= c) wso[S::WS_RN + c * (c + 1) / 2 + r] = Rq[t2];
                }
                if (l < wneq) wso[S::WS_RDI + l] = rdq[l];
                {   // inverse of the equality block of RN (lane c: column c by back substitution), for the solve kernel's
                    // thread-parallel "solves with RN"
                    __syncwarp();
                    double xr[F::NEQ];
#pragma unroll
                    for (int r = 0; r < F::NEQ; ++r) xr[r] = 0.0;
                    const int c = l < wneq ? l : 0;
#pragma unroll
                    for (int r = F::NEQ - 1; r >= 0; --r) {
                        double sacc = r == c ? -1.0 : 0.0;
#pragma unroll
                        for (int jj = r + 1; jj < F::NEQ; ++jj) if (jj <= c) sacc = fma(Rq[jj * F::NEQ + r], xr[jj], sacc);   // (columns > c: stale data)
                        xr[r] = r <= c ? -sacc * rdq[r < wneq ? r : 0] : 0.0;
                    }
                    if (l < wneq)
#pragma unroll
                        for (int r = 0; r < F::NEQ; ++r) if (r <= l) wso[S::WS_RI + l * (l + 1) / 2 + r] = xr[r];
                }
                wso[S::WS_U + i0] = ua;
                if (has1) wso[S::WS_U + i1] = ub;
                if (l == 0) wso[S::WS_FLAG] = flag;
            }
        }
        __syncthreads();
        if (tk.host) {                                         // hand the tick to the solve server; wait for the next one
            __threadfence();
            if (t == 0) { tick_stamp(tk, 2); st_release_gpu(tk.dev + 0, tick); }
            base = -(long long)gridDim.x * F::FPC;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Certificate kernel (shapes with P::SPLIT_FACTOR): the scaled KKT residual of SURVEY.md 8(c) for every level of
// every solved problem, from the record and the block the solve kernel exported (x, proximal centre, multipliers,
// working set) -- stationarity of the regularised, proximal-shifted level problem in x-space with the original
// task rows and the rebuilt constraint rows, primal feasibility of every row, complementarity and multiplier signs.
// It is regular, branch-poor code; inside the solve kernel it was 1/4 of the executed code bytes and evicted the
// active-set loop from the instruction caches of the SM (measured: +27 % solves/s without it, profiles/README.md).
// One CTA of CERT_THREADS threads per (problem, level); thread per constraint row, then thread per variable.
// ------------------------------------------------------------------------------------------
constexpr int CERT_THREADS = 128;
__device__ __forceinline__ void cert_max(double* slot, double v)     // v >= 0 (or NaN, which must win): order as integers
{
    atomicMax(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)__double_as_longlong(v));
}

// Resident certify server: copy the finished output record (x, tau, trailer) to pinned host memory, then the tick.
__device__ __forceinline__ void tick_publish(const Tick& tk, const unsigned char* out, int out_bytes, uint32_t tick)
{
    __threadfence();
    __syncthreads();
    for (int i = threadIdx.x; i < out_bytes / 8; i += blockDim.x) tk.host_out[i] = reinterpret_cast<const double*>(out)[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { tick_stamp(tk, 6); st_release_sys(tk.host + 1, tick); }
}

// One CTA per problem: the record is staged in shared memory once (coalesced), threads 0-63 certify level 0 and threads
// 64-127 level 1 side by side (same code, same barriers); inside a level a thread per constraint row, then a thread
// per variable.
template <class P>
__global__ void __launch_bounds__(CERT_THREADS)
qp_certify_kernel(const double* __restrict__ recs, unsigned char* __restrict__ out, const double* __restrict__ ws,
                  long long batch, Params prm, Tick tk)
{
    using S = Slab<P>;
    constexpr int N = P::N, NROWS = P::NROWS, T = 64;
    constexpr int OUT_BYTES = 8 * (N + P::NA) + 32;
    constexpr int MDP = P::MD_MAX + 2;
    double* const g = reinterpret_cast<double*>(g_smem);       // the record (P::REC doubles)
    __shared__ double xs[2][S::VEC], xps[2][S::VEC], ax[2][MDP], bx[2][MDP], eo[2][8], yrow[2][NROWS], mx[2][8];
    __shared__ signed char sg_row[2][NROWS];
    __shared__ int act[2][S::KP + 2];
    const int level = threadIdx.x >> 6, t = threadIdx.x & 63;
    __shared__ uint32_t s_go, s_tick;
    uint32_t tick = tk.seq0;
#pragma unroll 1
    for (long long idx = blockIdx.x; idx < batch; idx += gridDim.x) {
        if (tk.host) {
            // resident: problem 0 once the solve server is done with it; afterwards the output record and the
            // completed tick go back to pinned host memory
            if (threadIdx.x == 0) { uint32_t nt = tick; s_go = tick_wait_dev(tk, 1, 3, nt) ? 1u : 0u; s_tick = nt; }
            __syncthreads();
            const uint32_t go = s_go;
            tick = s_tick;
            __syncthreads();
            if (threadIdx.x == 0 && !go) st_release_sys(tk.host + 3, 0u);                  // the chain is down
            if (!go) break;
            if (threadIdx.x == 0) tick_stamp(tk, 5);
            idx = 0;
        }
        uint32_t* tr = reinterpret_cast<uint32_t*>(out + idx * (size_t)OUT_BYTES + 8 * (N + P::NA));
        if (!tk.host && (int)tr[0] != QPPVM_STATUS_OK) continue;          // failed solves keep kkt = +inf (CTA-uniform)
        if (tk.host && (int)tr[0] != QPPVM_STATUS_OK) {
            tick_publish(tk, out, OUT_BYTES, tick);
            idx = -(long long)gridDim.x;
            continue;
        }
        const double* gr = recs + idx * (size_t)P::REC;
        for (int i = threadIdx.x; i < P::REC / 2; i += CERT_THREADS)
            reinterpret_cast<double2*>(g)[i] = reinterpret_cast<const double2*>(gr)[i];
        const double* cb = ws + idx * (size_t)S::WS + level * S::WS_LEVEL;
        const int* ci = reinterpret_cast<const int*>(cb + S::C_INT);
        const int k = ci[0];
        for (int i = t; i < N; i += T) { xs[level][i] = cb[S::C_X + i]; xps[level][i] = cb[S::C_XP + i]; }
        for (int r = t; r < NROWS; r += T) { yrow[level][r] = 0.0; sg_row[level][r] = 0; }
        if (t < 8) { mx[level][t] = 0.0; eo[level][t] = t < QPPVM_M0 ? cb[S::C_EOPT + t] : 0.0; }
        __syncthreads();
        if (t < k) {
            const int v = ci[1 + t], row = v & 0xff;
            act[level][t] = row; yrow[level][row] = cb[S::C_Y + t]; sg_row[level][row] = (signed char)(v >> 8);
        }
        __syncthreads();
        const double* const x = xs[level];
        const double eps = P::regularised(level) ? prm.eps_reg : 0.0;
        // ---- constraint rows: primal feasibility, complementarity, multiplier signs
        double rprim = 0.0, rcomp = 0.0, cxmax = 0.0, ymax = 0.0;
        for (int r = t; r < NROWS; r += T) {
            double val, lo, hi;
            if (!P::row_value(g, level, r, x, eo[level], val, lo, hi)) continue;
            const double y = yrow[level][r];
            const int sg = sg_row[level][r];
            cxmax = fmax(cxmax, fabs(val)); ymax = fmax(ymax, fabs(y));
            if (sg != 0 && !(sg & 1)) rprim = fmax(rprim, fabs(val - lo));                 // equality in the working set
            else {
                rprim = fmax(rprim, fmax(0.0, fmax(lo - val, val - hi)));
                if (sg != 0) {
                    rcomp = fmax(rcomp, y > 0.0 ? y * fabs(val - lo) : -y * fabs(hi - val));
                    if (sg * y < 0.0) rcomp = fmax(rcomp, fabs(y));                         // wrong-signed multiplier
                }
            }
            if (!(val == val)) rprim = val;                                                // NaN must surface
        }
        // ---- task rows: (A x)_r and b_r
        const int md = P::task_rows(level);
        for (int r = t; r < md; r += T) {
            double s0 = 0.0, s1 = 0.0;
            int j = 0;
#pragma unroll 1
            for (; j + 1 < P::NB; j += 2) {
                s0 = fma(P::task_coef(g, level, r, j, prm), x[j], s0);
                s1 = fma(P::task_coef(g, level, r, j + 1, prm), x[j + 1], s1);
            }
            if (j < P::NB) s0 = fma(P::task_coef(g, level, r, j, prm), x[j], s0);
            ax[level][r] = s0 + s1; bx[level][r] = P::task_rhs(g, level, r, prm);
        }
        __syncthreads();
        // ---- stationarity, a variable per thread: H x + g - sum_c y_c a_c
        double rs = 0.0, gmax = 0.0, hxmax = 0.0, xmax = 0.0;
        for (int j = t; j < N; j += T) {
            double dgv, dbv;
            P::task_diag(g, level, j, dgv, dbv, prm);
            double hx = (dgv + eps) * x[j], gg = -dgv * dbv - eps * xps[level][j];
            if (j < P::NB)
#pragma unroll 1
                for (int r = 0; r < md; ++r) { const double a = P::task_coef(g, level, r, j, prm); hx = fma(a, ax[level][r], hx); gg = fma(-a, bx[level][r], gg); }
            double cy = 0.0;
#pragma unroll 1
            for (int c = 0; c < k; ++c) { const int row = act[level][c]; cy = fma(yrow[level][row], P::row_coef(g, row, j), cy); }
            const double stn = hx + gg - cy;
            rs = fmax(rs, fabs(stn)); gmax = fmax(gmax, fabs(gg)); hxmax = fmax(hxmax, fabs(hx)); xmax = fmax(xmax, fabs(x[j]));
            if (!(stn == stn)) rs = stn;
        }
        double* const m = mx[level];
        cert_max(m + 0, rs); cert_max(m + 1, gmax); cert_max(m + 2, hxmax); cert_max(m + 3, xmax);
        cert_max(m + 4, rprim); cert_max(m + 5, rcomp); cert_max(m + 6, cxmax); cert_max(m + 7, ymax);
        __syncthreads();
        if (t == 0) {
            const double a = m[0] / fmax(1.0, fmax(m[1], m[2]));
            const double b = m[4] / fmax(1.0, fmax(m[3], m[6]));
            const double c = m[5] / (fmax(1.0, m[7]) * fmax(1.0, m[6]));
            double kv = fmax(a, fmax(b, c));
            if (!(a == a) || !(b == b) || !(c == c)) kv = a + b + c;                       // NaN
            tr[6 + level] = __float_as_uint((float)kv);
        }
        __syncthreads();
        if (tk.host) {
            tick_publish(tk, out, OUT_BYTES, tick);
            idx = -(long long)gridDim.x;
        }
    }
}

}  // namespace qppvm
