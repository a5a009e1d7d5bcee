/*
 * qppvm_b200.h — C-ABI of the B200-native whole-body QP hot path.
 *
 * This is the drop-in boundary under the two XBotCore RT plugins of the
 * reference.  One call to qppvm_solve_batch() does, for every record of a
 * batch, what one tick of the reference does between "model is updated" and
 * "torques are written":
 *
 *   ForceAcc kind  (QPPVM_KIND_FORCEACC)
 *     replaces  _autostack->update()            ref:src/ForceAcc.cpp:184
 *               _solver->solve(_x)              ref:src/ForceAcc.cpp:188-193
 *               qddot/wrench getValue           ref:src/ForceAcc.cpp:196-201
 *               tau = ID(qddot) - sum J^T w     ref:src/ForceAcc.cpp:206-219
 *     problem structure (variables, tasks, bounds, eps=1e4)
 *                                               ref:src/ForceAcc.cpp:58-137
 *
 *   Torque kind  (QPPVM_KIND_TORQUE)
 *     replaces  torque-limit shift by h         ref:src/QPPVMPlugin.cpp:203-205
 *               _autostack->update(_q)          ref:src/QPPVMPlugin.cpp:226
 *               _solver->solve(_tau_d)          ref:src/QPPVMPlugin.cpp:246-249
 *               _tau_d += _h                    ref:src/QPPVMPlugin.cpp:256
 *     problem structure (tasks, gains, eps=1.0) ref:src/QPPVMPlugin.cpp:112-188
 *
 * Plain C: POD structs, raw pointers and sizes.  No C++/torch types cross it.
 * One handle per calling thread (the reference's single-RT-thread contract).
 * Every entry point returns 0 on success (== reference's solve()==true for the
 * call as a whole; per-problem solver status is in the output trailer).
 */
#ifndef QPPVM_B200_H_
#define QPPVM_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- problem kinds -------------------------------------------------------- */
#define QPPVM_KIND_TORQUE    0  /* x = tau            (ref:src/QPPVMPlugin.cpp)   */
#define QPPVM_KIND_FORCEACC  1  /* x = [qddot ; f_c]  (ref:src/ForceAcc.cpp:63-70) */

/* ---- optional constraint families (north_star; OpenSoT definitions) ------- */
#define QPPVM_FLAG_FRICTION_CONES  1  /* 5-row linearised pyramid per contact    */
#define QPPVM_FLAG_TORQUE_LIMITS   2  /* tau_min <= M_a qdd + h_a - J^T f <= max  */
#define QPPVM_FLAG_FULL_WRENCH     4  /* 6 variables per contact (force + torque): "put 6 for full wrench",
                                         ref:src/ForceAcc.cpp:67; the wrench bounds are then six real rows per contact */
#define QPPVM_FLAG_COM_TASK       32  /* FORCEACC: the centroidal force task the reference constructs (OpenSoT tasks::force::CoM,
                                         ref:src/ForceAcc.cpp:103) joins level 1: postural + contact Cartesian + CoM.  Its six
                                         rows act on the contact force / wrench variables, which makes those columns dense in
                                         the level-1 Hessian (general dense path of the kernels)                          */
/* TORQUE kind: the tasks / constraints the reference instantiates next to its stack */
#define QPPVM_FLAG_JOINT_LIMITS    8  /* torque-domain joint limits (OpenSoT constraints::torque::JointLimits,
                                         ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:70, ref:src/QPPVMPlugin.cpp:169-171):
                                         simple bounds on tau, intersected with the shifted torque limits            */
#define QPPVM_FLAG_ELBOW_TASKS    16  /* level 1 = elbow_left + elbow_right (two 3-row Cartesian impedance tasks,
                                         ref:src/QPPVMPlugin.cpp:154-166) instead of the joint impedance task: the
                                         alternative stack of ref:src/QPPVMPlugin.cpp:177-178                         */

/* ---- per-problem solver status (reference: bool from solve()) ------------- */
#define QPPVM_STATUS_OK          0
#define QPPVM_STATUS_MAX_ITER    1  /* nWSR exhausted                            */
#define QPPVM_STATUS_INFEASIBLE  2
#define QPPVM_STATUS_NUMERIC     3  /* active-set overflow / non-finite data     */

/* ---- API return codes ------------------------------------------------------ */
#define QPPVM_OK                 0
#define QPPVM_ERR_ARG            1
#define QPPVM_ERR_UNSUPPORTED    2  /* no kernel instantiated for these dims     */
#define QPPVM_ERR_CUDA           3
#define QPPVM_ERR_NO_DEVICE      4

#define QPPVM_QPOASES_EPS        2.221e-16         /* qpOASES EPS                */
#define QPPVM_QPOASES_EPS_REG    (1.0e3 * QPPVM_QPOASES_EPS) /* default epsRegularisation */
#define QPPVM_INFTY              1.0e20            /* qpOASES INFTY              */
#define QPPVM_M0                 6                 /* rows of the level-0 task   */

typedef struct qppvm_desc {
    int32_t kind;               /* QPPVM_KIND_*                                          */
    int32_t n_a;                /* actuated joints                                       */
    int32_t n_contacts;         /* FORCEACC: contacts c; TORQUE: must be 2 (two hands)   */
    int32_t flags;              /* QPPVM_FLAG_*  (1, 2, 4, 32: FORCEACC; 8, 16: TORQUE)  */
    double  eps_regularisation; /* QPOases_sot ctor arg: 1e4 (ForceAcc.cpp:137) / 1.0    */
    int32_t n_reg_steps;        /* qpOASES numRegularisationSteps (MPC option set: 1)    */
    int32_t max_iter;           /* working-set changes allowed per level (nWSR): 132     */
    int32_t device;             /* CUDA ordinal                                          */
    int32_t postural_actuated_only; /* FORCEACC: 1 = the Postural task has no rows for the 6 floating-base velocities
                                   (later OpenSoT versions); 0 = A = [I 0] on all n_v rows    */
    /* Upstream semantics that varied across OpenSoT versions ("make it a parameter").  FORCEACC kind.
     * 0.0 means "not set" and is read as 1.0, so a zero-initialised tail reproduces the consistent reading. */
    double  lambda_solver;      /* QPOases_sot: g = -lambda A^T W b (A.2): multiplies every task right-hand side */
    double  task_weight[3];     /* W = w I per task: waist (level 0), postural, contact-link Cartesian (level 1) */
} qppvm_desc;

/*
 * Record layout (one problem), all FP64, offsets in doubles, problem-major
 * contiguous; record stride is padded to an even number of doubles (16 B).
 *
 * FORCEACC  (n_v = n_a + 6, n_x = n_v + w c with w = 3 force components per contact, 6 with QPPVM_FLAG_FULL_WRENCH):
 *   J_waist  6 x n_v row-major          level-0 Cartesian task Jacobian   (ForceAcc.cpp:118-122)
 *   J_c      c x 6 x n_v                contact-link Jacobians            (ForceAcc.cpp:83-89, 208)
 *   M        packed lower, row-major    n_v(n_v+1)/2                      (DynamicFeasibility, ID)
 *   h        n_v                        nonlinear term
 *   Jdqd     6(1+c)                     Jdot*qdot, waist first
 *   rhs      6(1+c) + n_v               a_ref + l2*edot + l*e per Cartesian task, then postural
 *   tau_min, tau_max   2 n_a            only with QPPVM_FLAG_TORQUE_LIMITS
 *   cone     c x (R 3x3 row-major, mu)  only with QPPVM_FLAG_FRICTION_CONES
 *   f_lb,f_ub  c x (lb w, ub w)         force / wrench box                 (ForceAcc.cpp:74-76)
 *   A_com 6 x (w c) row-major, b_com 6  only with QPPVM_FLAG_COM_TASK: centroidal dynamics on the contact wrenches,
 *                                       [sum f_i ; sum (p_i - c) x f_i (+ tau_i)] = [m (a_ref - g) ; Ldot_ref]   (ForceAcc.cpp:103)
 *
 * TORQUE  (n_v = n_x = n_a):
 *   J_ee     2 x 6 x n  (right hand first: stack order ee_right + ee_left, QPPVMPlugin.cpp:177)
 *   M        packed lower n(n+1)/2
 *   h        n
 *   F_ee     2 x 6      K e + D edot per hand (spring+damper wrench)
 *   tau_j    n          K (q_ref - q) + D (-qdot)                          (QPPVMPlugin.cpp:105-118)
 *   tau_min_const, tau_max_const  2 n  (before the -h shift of QPPVMPlugin.cpp:203-204)
 *   jl_min, jl_max   2 n    only with QPPVM_FLAG_JOINT_LIMITS: k (q_min - q) - d qdot, k (q_max - q) - d qdot
 *                           (OpenSoT torque::JointLimits::update; gains setGains(k, d), QPPVMPlugin.cpp:170);
 *                           NOT shifted by h: they bound the solver's variable directly
 *   J_elbow  2 x 6 x n      only with QPPVM_FLAG_ELBOW_TASKS (left elbow first: stack order elbow_left + elbow_right,
 *                           QPPVMPlugin.cpp:178), then F_elbow 2 x 6 (K e + D edot per elbow)
 */
typedef struct qppvm_layout {
    int32_t n_a, n_v, n_c, n_x;
    int32_t n_rows;        /* constraint rows incl. level-1 optimality rows (mask width)   */
    int32_t row_dyn;       /* first dyn-feas row (6 rows)          ; -1 if absent          */
    int32_t row_box;       /* first force-box row (6 per contact) / simple bounds (TORQUE) */
    int32_t row_cone;      /* first friction row (5 per contact)   ; -1 if absent          */
    int32_t row_tau;       /* first torque-limit row (n_a rows)    ; -1 if absent          */
    int32_t row_opt;       /* first level-1 optimality row (QPPVM_M0 rows)                 */
    int32_t off_jwaist, off_jc, off_M, off_h, off_jdqd, off_rhs;
    int32_t off_taulim, off_cone, off_fbox;          /* -1 if absent                        */
    int32_t off_fee, off_tauj;                       /* TORQUE kind only, else -1          */
    int32_t rec_doubles;   /* record stride in doubles (even)                               */
    int32_t out_bytes;     /* output stride: 8 (n_x + n_a) + 32                             */
    int32_t diag_doubles;  /* diagnostic stride: n_x (x0) + 2 n_rows (y0,y1) + QPPVM_M0     */
    int32_t off_jlim, off_jelbow, off_felbow;        /* TORQUE kind with the flags above, else -1 */
    int32_t off_com;                                 /* FORCEACC with QPPVM_FLAG_COM_TASK: A_com | b_com, else -1 */
} qppvm_layout;

/* Output trailer that follows x[n_x], tau[n_a] in every output record (32 B). */
typedef struct qppvm_trailer {
    int32_t  status;       /* QPPVM_STATUS_*                                   */
    int32_t  iters;        /* working-set changes: level 0 | (level 1 << 16)   */
    uint32_t active[4];    /* bit r set <=> constraint row r active at level 1 */
    float    kkt[2];       /* scaled KKT residual per level      */
} qppvm_trailer;

typedef struct qppvm_handle qppvm_handle;

/* Pure host arithmetic; usable without a GPU.  Returns QPPVM_ERR_ARG on bad dims. */
int qppvm_get_layout(const qppvm_desc* desc, qppvm_layout* out);

/* Creates the solver for one problem shape.  Fails (no CPU fallback) when there is
 * no CUDA device or no kernel instantiated for the shape. */
int qppvm_create(const qppvm_desc* desc, qppvm_handle** out);
int qppvm_destroy(qppvm_handle* h);
/* Message of the last failure on this handle (NULL handle: last create failure). */
const char* qppvm_last_error(const qppvm_handle* h);

/* Batched solve, device pointers, asynchronous on `cuda_stream` (cudaStream_t or NULL).
 * records: B * rec_doubles doubles.  out: B * out_bytes bytes. */
int qppvm_solve_batch(qppvm_handle* h, const double* records_dev, void* out_dev,
                      int64_t batch, void* cuda_stream);
/* Same plus the diagnostic block (level-0 solution and the multipliers of both levels). */
int qppvm_solve_batch_diag(qppvm_handle* h, const double* records_dev, void* out_dev,
                           double* diag_dev, int64_t batch, void* cuda_stream);
/* Hot-started batched solve.  The reference keeps ONE QPOases_sot alive across control ticks
 * (ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:64, built once ref:src/QPPVMPlugin.cpp:188 / ref:src/ForceAcc.cpp:135-137,
 * called every tick ref:src/QPPVMPlugin.cpp:246 / ref:src/ForceAcc.cpp:189), so qpOASES starts every tick from the
 * previous tick's working set.  `warm_dev` is that state for a batch: QPPVM_WARM_WORDS uint32 per problem (active-row
 * masks of level 0 | level 1, row ids as in qppvm_layout), read at the start of each problem and overwritten with the
 * working set it converged to; all-zero words = cold start.  The result is the same unique minimiser either way. */
#define QPPVM_WARM_WORDS 8
int qppvm_solve_batch_warm(qppvm_handle* h, const double* records_dev, void* out_dev, uint32_t* warm_dev,
                           int64_t batch, void* cuda_stream);
/* Batched solve with HOST buffers: chunked H2D / solve / D2H overlapped on internal
 * streams; returns after the outputs are in `out_host`. */
int qppvm_solve_batch_host(qppvm_handle* h, const double* records_host, void* out_host,
                           int64_t batch);
/* Pipelined form of qppvm_solve_batch_host: enqueues the chunked H2D / solve / D2H work and returns; consecutive
 * calls overlap (the copies of batch i+1 run under the solve of batch i).  `records_host` must stay valid and
 * `out_host` must not be read until qppvm_host_sync() returns; use pinned buffers.  Calls on one handle are
 * executed in order. */
int qppvm_solve_batch_host_async(qppvm_handle* h, const double* records_host, void* out_host, int64_t batch);
int qppvm_host_sync(qppvm_handle* h);
/* Latency mode: one record, host in / host out, synchronous (one control tick: ref:src/QPPVMPlugin.cpp:308-329,
 * ref:src/ForceAcc.cpp:167-253).  For the ForceAcc shapes no kernel is launched per tick: three resident one-CTA
 * kernels (prepare, solve, certify) poll a mailbox in pinned host memory.  Consecutive ticks hot-start from the previous
 * tick's working sets, as the reference's persistent QPOases_sot does (ref:include/QPPVM_RT_plugin/QPPVMPlugin.h:64);
 * qppvm_reset_warm() makes the next tick a cold start.  Environment: QPPVM_RESIDENT=0 falls back to one launch sequence
 * per tick, QPPVM_TICK_IDLE_US (default 20000) is how long the resident kernels wait for a tick before they leave. */
int qppvm_solve_one(qppvm_handle* h, const double* record_host, void* out_host);
int qppvm_reset_warm(qppvm_handle* h);
/* Device-clock stamps (ns, %globaltimer) of the last tick through the resident chain: tick seen by the prepare server,
 * record copied to device memory, prepare done, solve start, solve done, certify start, result published. */
int qppvm_tick_stamps(qppvm_handle* h, uint64_t* ns7);

/* ---- rigid-body front end: compact states -> records on the device ------------------
 * Replaces, for batched use, what the reference obtains on the CPU from XBot::ModelInterface after
 * model->update() (ref:src/ForceAcc.cpp:256-282, getJacobian :208, M and h inside DynamicFeasibility :109-114)
 * and the OpenSoT task right-hand sides.  ForceAcc kind only.
 * State layout (doubles): q n_a | qd n_a | R0 3x3 row-major | p0 3 | base twist (v0, w0) 6 | gains 4
 *                         (lambda, lambda2 factors: waist, postural/contacts) | waist orientation error 3 |
 *                         contact pose errors 6c | mu c | tau-limit scale n_a | waist position error 3
 *                         (errors = reference - current; the references are captured once, ref:src/ForceAcc.cpp:158-164,
 *                         waist position reference = initial - 0.1 z, :181).                                       */
typedef struct qppvm_robot {
    int32_t n_a;                    /* actuated joints; bodies = n_a + 1, body 0 = floating base            */
    const int32_t* parent;          /* [n_a + 1] parent body of each body, parent[0] = -1                   */
    const double* axis;             /* [n_a + 1][3] joint axis in the parent frame (unit)                   */
    const double* offset;           /* [n_a + 1][3] joint origin in the parent frame                        */
    const double* mass;             /* [n_a + 1]                                                            */
    const double* com;              /* [n_a + 1][3] centre of mass in the body frame                        */
    const double* inertia;          /* [n_a + 1][3] principal inertia about the COM (body axes)             */
    const double* q_home;           /* [n_a] postural reference                                             */
    const double* tau_max;          /* [n_a] effort limits (scaled per state by the tau-limit scale)        */
    const int32_t* contact_body;    /* [n_contacts] contact link bodies, in force-variable order            */
} qppvm_robot;
int qppvm_state_doubles(const qppvm_desc* desc);                    /* -1 on a bad description */
int qppvm_set_robot(qppvm_handle* h, const qppvm_robot* robot);     /* copies the tables to the device */
int qppvm_records_from_states(qppvm_handle* h, const double* states_dev, double* records_dev,
                              int64_t batch, void* cuda_stream);
/* states (host) -> records (device, never leave it) -> solve -> outputs (host); chunked and overlapped. */
int qppvm_solve_states_host(qppvm_handle* h, const double* states_host, void* out_host, int64_t batch);
/* Pipelined form (see qppvm_solve_batch_host_async); completed by qppvm_host_sync(). */
int qppvm_solve_states_host_async(qppvm_handle* h, const double* states_host, void* out_host, int64_t batch);

/* ---- command side: integrate the solved accelerations, closed-loop rollouts on the device ----
 * qppvm_integrate_states: states (device, in/out) advance by one period dt with the q-ddot of `out` (device, the block
 * qppvm_solve_batch wrote for the same states): q += dt*qd + dt^2/2*qdd, qd += dt*qdd, floating base likewise
 * (ref:src/ForceAcc.cpp:225-226).  A state whose solve failed is left as it is (ref:src/ForceAcc.cpp:189-193).
 * qppvm_rollout_states: `ticks` control periods of front end -> 2-level solve -> integrate, all on `stream`, nothing
 * crossing PCIe; `out` holds the last tick's solutions.  The task references are those the states were created with: the
 * stored errors shrink as the robot moves towards them.  Each tick of a state starts from the working sets of its previous
 * tick (the persistent QPOases_sot of the reference hot-starts the same way, ref:src/ForceAcc.cpp:135-137,189). */
int qppvm_integrate_states(qppvm_handle* h, double* states_dev, const void* out_dev, double dt, int64_t batch, void* stream);
/* The same plus the tick's records: the task errors stored in the states follow the motion (e <- e - dt v_link -
 * dt^2/2 a_link per task link), which is what keeps multi-tick rollouts closed-loop in the task references;
 * qppvm_rollout_states integrates this way. */
int qppvm_integrate_states_tracking(qppvm_handle* h, double* states_dev, const void* out_dev, const double* records_dev,
                                    double dt, int64_t batch, void* stream);
int qppvm_rollout_states(qppvm_handle* h, double* states_dev, void* out_dev, int ticks, double dt, int64_t batch, void* stream);

/* ---- the batch sharded over the GPUs of one box, single process ---------------------------------------
 * The path shards with no data-path collective (every state's cascade is independent: ref:src/QPPVMPlugin.cpp:246,
 * ref:src/ForceAcc.cpp:189): GPU r of G solves the contiguous block [r B / G, (r + 1) B / G).  `devices` = CUDA ordinals
 * (NULL: 0 .. n - 1), devices[0] is the root.  qppvm_multi_solve_batch: records and outputs on the ROOT GPU; the other
 * GPUs' blocks travel by grouped ncclSend / ncclRecv (ncclCommInitAll, NVLink / NVSwitch), in chunks, so that the scatter of
 * chunk i + 1 and the gather of chunk i - 1 overlap the solve of chunk i; outputs land directly in their place in
 * `out_root_dev`.  Synchronous.  qppvm_multi_solve_states: the same with compact STATES on the root GPU (after
 * qppvm_multi_set_robot): the states travel (1 KB instead of 11-18 KB per problem, so the root's NVLink egress is no longer
 * the bound) and every GPU runs the rigid-body front end on its chunk before the solve.  The *_host forms take host (pinned) buffers: every GPU moves its own block over its own
 * PCIe link, no GPU-to-GPU traffic.  Results are bitwise those of one GPU. */
typedef struct qppvm_multi qppvm_multi;
int qppvm_multi_create(const qppvm_desc* desc, const int32_t* devices, int n_devices, qppvm_multi** out);
int qppvm_multi_destroy(qppvm_multi* m);
const char* qppvm_multi_last_error(const qppvm_multi* m);
int qppvm_multi_devices(const qppvm_multi* m);
int qppvm_multi_set_robot(qppvm_multi* m, const qppvm_robot* robot);
int qppvm_multi_solve_batch(qppvm_multi* m, const double* records_root_dev, void* out_root_dev, int64_t batch);
int qppvm_multi_solve_states(qppvm_multi* m, const double* states_root_dev, void* out_root_dev, int64_t batch);
int qppvm_multi_solve_batch_host(qppvm_multi* m, const double* records_host, void* out_host, int64_t batch);
int qppvm_multi_solve_states_host(qppvm_multi* m, const double* states_host, void* out_host, int64_t batch);
int64_t qppvm_multi_kernel_launches(const qppvm_multi* m);
int64_t qppvm_multi_nccl_calls(const qppvm_multi* m);

/* Per-kernel device time (CUDA events on the launching stream around every launch of the prepare / solve / certify
 * kernels).  Returns in ms3 / launches3 the totals accumulated since the previous call (both may be NULL), forgets them,
 * and switches the recording on or off.  Meant for measurements: the events serialise nothing but cost a few us each. */
int qppvm_kernel_timing(qppvm_handle* h, int enable, double* ms3, int64_t* launches3);
/* The persistent grids of this handle leave `n_sms` SMs free (default 0): room for kernels that must run concurrently
 * with a long solve, e.g. NCCL's copy kernels on the root GPU of qppvm_multi_solve_batch. */
int qppvm_reserve_sms(qppvm_handle* h, int n_sms);
/* Number of kernel launches issued through this handle so far. */
int64_t qppvm_kernel_launches(const qppvm_handle* h);
/* Measures the FP64 FMA peak of the device (TFLOP/s) with a register-resident DFMA
 * kernel; the compute-roofline denominator (not in MEASURED_PEAKS.json). */
int qppvm_fp64_peak(qppvm_handle* h, double* tflops);
/* Shapes compiled into this library: writes up to `cap` QUADRUPLES (kind, n_a, n_contacts, flags), i.e. 4 * cap
 * int32, and returns the number of shapes available (call with quads = NULL or cap = 0 to size the buffer). */
int qppvm_supported_shapes(int32_t* quads, int cap);

#ifdef __cplusplus
}
#endif
#endif /* QPPVM_B200_H_ */
