// qppvm_capi.cu — C-ABI (include/qppvm_b200.h) over the sm_100a kernels.  Plain CUDA runtime;
// no torch, no CPU fallback: every solve entry point fails when there is no device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "qp_kernel.cuh"
#include <vector>
#include "rbd_kernel.cuh"

using namespace qppvm;

namespace {

char g_create_error[512] = "";

struct ShapeEntry {
    int kind, n_a, n_c, flags;
    const void* kernel;          // qp_solve_kernel<P, 64>: 64 threads (one CTA) per problem
    int slab_bytes;
    const void* factor_kernel;   // qp_factor_kernel<P> for shapes that factor in a separate launch, else null
    int ws_doubles;              // factor workspace per problem (doubles)
    int factor_threads, factor_pairs, factor_bytes;   // CTA size, (problem, level) pairs per CTA, dynamic smem
    const void* certify_kernel;  // qp_certify_kernel<P> (KKT certificate of the shapes above), else null
};

// Instantiated problem shapes (BASELINE.json configs).
template <class P>
constexpr const void* factor_entry()
{
    if constexpr (P::SPLIT_FACTOR) return (const void*)&qp_factor_kernel<P>;
    else return nullptr;
}
template <class P>
constexpr const void* certify_entry()
{
    if constexpr (P::SPLIT_FACTOR) return (const void*)&qp_certify_kernel<P>;
    else return nullptr;
}
template <class P>
constexpr ShapeEntry entry()
{
    return ShapeEntry{P::KIND, P::NA, P::NC, P::FLAGS,
                      (const void*)&qp_solve_kernel<P, 64>, Slab<P>::BYTES,
                      factor_entry<P>(), Slab<P>::WS,
                      FactorShape<P>::THREADS, FactorShape<P>::FPC, FactorShape<P>::BYTES, certify_entry<P>()};
}
// The instantiation list is generated from the build-time table shapes.def (one line per robot / constraint set).
const ShapeEntry g_shapes[] = {
#define QPPVM_SHAPE_FORCEACC(NA, NC, FLAGS) entry<ForceAcc<NA, NC, (FLAGS)>>(),
#define QPPVM_SHAPE_TORQUE(NA, FLAGS) entry<Torque<NA, (FLAGS)>>(),
#include "shapes.def"
#undef QPPVM_SHAPE_FORCEACC
#undef QPPVM_SHAPE_TORQUE
};
constexpr int N_SHAPES = sizeof(g_shapes) / sizeof(g_shapes[0]);

constexpr int HOST_STREAMS = 4;
constexpr int N_SLOTS = HOST_STREAMS + 2;  // launch slots: host / rollout streams, caller's stream, spare
// Problems per prepare-workspace pass on the caller's stream.  The workspace is 28.5 - 46.5 KB per problem, i.e.
// 0.9 - 1.5 GB per pass: it does NOT stay in the 126 MB L2 (ncu: the solve kernel reads it back from DRAM, < 8 % of the
// HBM bandwidth).  Passes are sized for throughput instead: tens of waves of resident CTAs, so that the idle tail of
// a pass (the last problems of a dynamic schedule) is a few per cent of it.
constexpr int64_t WS_CHUNK = 65536;
constexpr int64_t ROLL_LANE = 16384;       // states per lane of the on-device rollout (record scratch <= 4 x 0.3 GB)

}  // namespace

struct qppvm_handle {
    qppvm_desc desc;
    qppvm_layout L;
    const ShapeEntry* shape;
    const void* kernel;
    int team;                              // threads per problem (= CTA size)
    int sm_count, ctas_per_sm;
    int reserve_sms;                       // SMs left free by the grids (root of a multi-GPU run: room for NCCL's copy kernels)
    unsigned long long* counters;          // N_SLOTS device counters
    double* ws[N_SLOTS]; int64_t ws_cap[N_SLOTS];   // factor workspaces, one per launch slot, allocated on first use
    int factor_ctas_per_sm, certify_ctas_per_sm;
    int rowwise;                           // QPPVM_ROWWISE_EQUALITIES=1 at create (tests): see Params::rowwise
    cudaStream_t dev_stream; bool dev_used; cudaEvent_t ev_dev;   // last stream the caller's-stream slot ran on
    cudaStream_t streams[HOST_STREAMS];
    double* d_rec[HOST_STREAMS];
    unsigned char* d_out[HOST_STREAMS];
    int64_t chunk;                         // records per host-path chunk
    int64_t chunk_states;                  // states per chunk of the state front end (transfers are 10x smaller)
    double* d_state[HOST_STREAMS];         // host-path staging for the state front end
    RobotTables rob; RbdShape rsh; void* rob_blob; bool has_robot;
    bool rbd_v1; int rbd_smem, rbd_ctas_per_sm;   // front-end kernel choice and launch geometry (set by qppvm_set_robot)
    double* d_roll; int64_t roll_cap;
    uint32_t* d_roll_warm; int64_t roll_warm_cap;   // working sets carried from tick to tick of a rollout (QPPVM_WARM_WORDS per state)      // record scratch of the on-device rollout: HOST_STREAMS lanes of roll_cap
    cudaEvent_t ev_fork, ev_join[HOST_STREAMS];
    double* d_one_rec; unsigned char* d_one_out;
    double* h_one_rec; unsigned char* h_one_out;   // pinned staging for latency mode
    cudaStream_t one_stream;
    // latency mode: resident prepare / solve / certify servers (qp_kernel.cuh, struct Tick)
    uint32_t* tick_host; uint32_t* tick_dev; uint32_t* d_one_warm; double* one_ws;
    cudaStream_t tick_streams[3]; bool tick_running; uint32_t tick_seq; int resident; uint32_t idle_us;
    int64_t launches;
    // optional per-kernel timing (bench.py: share of the step per kernel): events around every launch of the three kernels
    bool timing; std::vector<cudaEvent_t>* tev;     // (kind, start, stop) triples flattened: kind in tkind
    std::vector<int>* tkind;
    char err[512];
};

namespace {

int fail(qppvm_handle* h, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h ? h->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

// Entry points run on the handle's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess; else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define ENTER(h)                                                                                         \
    DeviceGuard guard_((h)->desc.device);                                                                \
    if (!guard_.ok) return fail(h, QPPVM_ERR_CUDA, "cudaSetDevice(%d) failed", (h)->desc.device)

#define CU(h, call)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(h, QPPVM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// Records a start / stop event pair around one kernel launch when timing is on (kind: 0 prepare, 1 solve, 2 certify).
struct LaunchTimer {
    qppvm_handle* h; cudaStream_t st; cudaEvent_t e1 = nullptr;
    LaunchTimer(qppvm_handle* h_, int kind, cudaStream_t st_);
    ~LaunchTimer() { if (e1) cudaEventRecord(e1, st); }
};

// qppvm_desc -> kernel parameters (0.0 in the upstream-semantics fields means "not set": 1.0)
Params make_params(const qppvm_handle* h)
{
    Params p;
    memset(&p, 0, sizeof(p));
    p.eps_reg = h->desc.eps_regularisation * QPPVM_QPOASES_EPS_REG;
    p.n_reg_steps = h->desc.n_reg_steps; p.max_iter = h->desc.max_iter; p.rowwise = h->rowwise;
    p.lam = h->desc.lambda_solver != 0.0 ? h->desc.lambda_solver : 1.0;
    for (int k = 0; k < 3; ++k) p.sw[k] = sqrt(h->desc.task_weight[k] != 0.0 ? h->desc.task_weight[k] : 1.0);
    p.post_act_only = h->desc.postural_actuated_only != 0;
    return p;
}

int launch(qppvm_handle* h, const double* rec, void* out, double* diag, int64_t batch,
           cudaStream_t st, int slot, bool dynamic = true, uint32_t* warm = nullptr)
{
    if (batch <= 0) return QPPVM_OK;
    unsigned long long* counter = dynamic ? h->counters + slot : nullptr;   // null: static round-robin schedule
    const int sms = h->sm_count - h->reserve_sms;
    const long long cap = (long long)sms * h->ctas_per_sm;
    const Params prm = make_params(h);
    const bool split = h->shape->factor_kernel != nullptr;
    Tick no_tick;
    memset(&no_tick, 0, sizeof(no_tick));
    const int64_t pass = split ? h->ws_cap[slot] : batch;   // problems per prepare-workspace pass (allocated at create)
    for (int64_t c0 = 0; c0 < batch; c0 += pass) {
        long long b = batch - c0 < pass ? batch - c0 : pass;
        const double* r = rec + c0 * (size_t)h->L.rec_doubles;
        unsigned char* o = (unsigned char*)out + c0 * (size_t)h->L.out_bytes;
        double* dgp = diag ? diag + c0 * (size_t)h->L.diag_doubles : nullptr;
        double* ws = split ? h->ws[slot] : nullptr;
        if (split) {
            const long long fcap = (long long)sms * h->factor_ctas_per_sm;
            const long long fneed = (2 * b + h->shape->factor_pairs - 1) / h->shape->factor_pairs;
            const int fgrid = (int)(fneed < fcap ? fneed : fcap);
            void* fargs[] = {(void*)&r, (void*)&ws, (void*)&b, (void*)&prm, (void*)&counter, (void*)&no_tick};   // also resets the counter
            {
                LaunchTimer tm(h, 0, st);
                CU(h, cudaLaunchKernel(h->shape->factor_kernel, dim3(fgrid), dim3(h->shape->factor_threads), fargs,
                                       (size_t)h->shape->factor_bytes, st));
            }
            h->launches += 1;
        }
        if (counter && !split) CU(h, cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
        const int grid = (int)(b < cap ? b : cap);
        uint32_t* wm = warm ? warm + c0 * 8 : nullptr;
        void* args[] = {(void*)&r, (void*)&o, (void*)&dgp, (void*)&b, (void*)&prm, (void*)&counter, (void*)&ws, (void*)&wm, (void*)&no_tick};
        {
            LaunchTimer tm(h, 1, st);
            CU(h, cudaLaunchKernel(h->kernel, dim3(grid), dim3(h->team), args, (size_t)h->shape->slab_bytes, st));
        }
        h->launches += 1;
        if (split) {                                           // KKT certificate of the pass (reads the blocks the solve exported)
            const long long ccap = (long long)sms * h->certify_ctas_per_sm;
            const int cgrid = (int)(b < ccap ? b : ccap);
            const double* cws = ws;
            void* cargs[] = {(void*)&r, (void*)&o, (void*)&cws, (void*)&b, (void*)&prm, (void*)&no_tick};
            LaunchTimer tm(h, 2, st);
            CU(h, cudaLaunchKernel(h->shape->certify_kernel, dim3(cgrid), dim3(CERT_THREADS), cargs,
                                   sizeof(double) * (size_t)h->L.rec_doubles, st));
            h->launches += 1;
        }
    }
    return QPPVM_OK;
}

// Device-pointer entry points.  The caller's-stream slot (work counter, prepare workspace) is reused from call to
// call: every call leaves an event behind on the stream it used, and a call on a different stream waits for that event
// first -- the previous stream's handle itself is never touched again (the caller may have destroyed it).
int solve_on_caller_stream(qppvm_handle* h, const double* rec, void* out, double* diag, uint32_t* warm, bool want_warm,
                           int64_t batch, cudaStream_t st)
{
    if (!h) return QPPVM_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!rec || !out || (want_warm && !warm)))) return fail(h, QPPVM_ERR_ARG, "bad batch arguments");
    if (((uintptr_t)rec & 15) || ((uintptr_t)out & 7) || ((uintptr_t)diag & 7) || ((uintptr_t)warm & 3))
        return fail(h, QPPVM_ERR_ARG, "records must be 16-byte aligned (TMA bulk copy), outputs 8-byte, warm-start words 4-byte aligned");
    ENTER(h);
    if (h->dev_used && st != h->dev_stream) CU(h, cudaStreamWaitEvent(st, h->ev_dev, 0));
    h->dev_stream = st; h->dev_used = true;
    const int rc = launch(h, rec, out, diag, batch, st, HOST_STREAMS, true, warm);
    if (rc) return rc;
    CU(h, cudaEventRecord(h->ev_dev, st));
    return QPPVM_OK;
}

__global__ void fp64_peak_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---- latency mode: resident servers ------------------------------------------------------------------
void stop_servers(qppvm_handle* h)
{
    if (!h->tick_running) return;
    __atomic_store_n(&h->tick_host[2], 1u, __ATOMIC_RELEASE);              // quit request (the prepare server polls it)
    for (int i = 0; i < 3; ++i) cudaStreamSynchronize(h->tick_streams[i]);
    h->tick_running = false;
}

int start_servers(qppvm_handle* h)
{
    // previous servers (if any) have left: host[3] == 0 is the last thing the chain writes
    for (int i = 0; i < 3; ++i) CU(h, cudaStreamSynchronize(h->tick_streams[i]));
    const uint32_t done = h->tick_host[1];
    const uint32_t init[4] = {done, done, 0u, 0u};
    CU(h, cudaMemcpy(h->tick_dev, init, sizeof(init), cudaMemcpyHostToDevice));
    h->tick_host[2] = 0u;
    __atomic_store_n(&h->tick_host[3], 1u, __ATOMIC_RELEASE);
    Tick tk;
    tk.host = h->tick_host; tk.dev = h->tick_dev; tk.host_rec = h->h_one_rec; tk.dev_rec = h->d_one_rec;
    tk.host_out = reinterpret_cast<double*>(h->h_one_out); tk.seq0 = done; tk.idle_us = h->idle_us;
    Params prm = make_params(h);
    const double* rec = h->d_one_rec; unsigned char* out = h->d_one_out; double* dgp = nullptr;
    double* ws = h->ws[HOST_STREAMS + 1]; const double* cws = ws;
    long long b = 1; unsigned long long* counter = nullptr; uint32_t* wm = h->d_one_warm;
    void* fargs[] = {(void*)&rec, (void*)&ws, (void*)&b, (void*)&prm, (void*)&counter, (void*)&tk};
    CU(h, cudaLaunchKernel(h->shape->factor_kernel, dim3(1), dim3(h->shape->factor_threads), fargs, (size_t)h->shape->factor_bytes, h->tick_streams[0]));
    void* args[] = {(void*)&rec, (void*)&out, (void*)&dgp, (void*)&b, (void*)&prm, (void*)&counter, (void*)&ws, (void*)&wm, (void*)&tk};
    CU(h, cudaLaunchKernel(h->kernel, dim3(1), dim3(h->team), args, (size_t)h->shape->slab_bytes, h->tick_streams[1]));
    void* cargs[] = {(void*)&rec, (void*)&out, (void*)&cws, (void*)&b, (void*)&prm, (void*)&tk};
    CU(h, cudaLaunchKernel(h->shape->certify_kernel, dim3(1), dim3(CERT_THREADS), cargs, sizeof(double) * (size_t)h->L.rec_doubles, h->tick_streams[2]));
    h->launches += 3;
    h->tick_running = true;
    return QPPVM_OK;
}

LaunchTimer::LaunchTimer(qppvm_handle* h_, int kind, cudaStream_t st_) : h(h_), st(st_)
{
    if (!h->timing) return;
    cudaEvent_t e0 = nullptr;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e1 = nullptr; return; }
    cudaEventRecord(e0, st);
    h->tev->push_back(e0); h->tev->push_back(e1); h->tkind->push_back(kind);
}

}  // namespace

extern "C" {

int qppvm_get_layout(const qppvm_desc* d, qppvm_layout* L)
{
    if (!d || !L) return QPPVM_ERR_ARG;
    memset(L, 0, sizeof(*L));
    L->off_jlim = L->off_jelbow = L->off_felbow = L->off_com = -1;
    if (d->n_a < 1 || d->n_a > 58) return QPPVM_ERR_ARG;
    int off = 0, row = 0;
    if (d->kind == QPPVM_KIND_FORCEACC) {
        const int c = d->n_contacts;
        if (c < 1 || c > 4) return QPPVM_ERR_ARG;
        const bool cones = d->flags & QPPVM_FLAG_FRICTION_CONES, tl = d->flags & QPPVM_FLAG_TORQUE_LIMITS;
        const int nv = d->n_a + 6;
        const int wd = (d->flags & QPPVM_FLAG_FULL_WRENCH) ? 6 : 3;
        L->n_a = d->n_a; L->n_v = nv; L->n_c = c; L->n_x = nv + wd * c;
        if (L->n_x > 64) return QPPVM_ERR_ARG;
        L->row_dyn = row; row += 6;
        L->row_box = row; row += 6 * c;
        L->row_cone = cones ? row : -1; if (cones) row += 5 * c;
        L->row_tau = tl ? row : -1; if (tl) row += d->n_a;
        L->row_opt = row; row += QPPVM_M0;
        L->off_jwaist = off; off += 6 * nv;
        L->off_jc = off; off += c * 6 * nv;
        L->off_M = off; off += nv * (nv + 1) / 2;
        L->off_h = off; off += nv;
        L->off_jdqd = off; off += 6 * (1 + c);
        L->off_rhs = off; off += 6 * (1 + c) + nv;
        L->off_taulim = tl ? off : -1; if (tl) off += 2 * d->n_a;
        L->off_cone = cones ? off : -1; if (cones) off += 10 * c;
        L->off_fbox = off; off += 2 * wd * c;
        if (d->flags & QPPVM_FLAG_COM_TASK) { L->off_com = off; off += 6 * wd * c + 6; }
        if (d->flags & ~(QPPVM_FLAG_FRICTION_CONES | QPPVM_FLAG_TORQUE_LIMITS | QPPVM_FLAG_FULL_WRENCH | QPPVM_FLAG_COM_TASK)) return QPPVM_ERR_ARG;
        L->off_fee = L->off_tauj = -1;
    } else if (d->kind == QPPVM_KIND_TORQUE) {
        if (d->n_contacts != 2 || (d->flags & ~(QPPVM_FLAG_JOINT_LIMITS | QPPVM_FLAG_ELBOW_TASKS))) return QPPVM_ERR_ARG;
        const int n = d->n_a;
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
    } else return QPPVM_ERR_ARG;
    if (row > 128) return QPPVM_ERR_ARG;
    L->n_rows = row;
    L->rec_doubles = off + (off & 1);
    L->out_bytes = 8 * (L->n_x + L->n_a) + 32;
    L->diag_doubles = L->n_x + 2 * row + QPPVM_M0;
    return QPPVM_OK;
}

int qppvm_supported_shapes(int32_t* t, int cap)
{
    for (int i = 0; t && i < N_SHAPES && i < cap; ++i) {
        t[4 * i] = g_shapes[i].kind; t[4 * i + 1] = g_shapes[i].n_a; t[4 * i + 2] = g_shapes[i].n_c; t[4 * i + 3] = g_shapes[i].flags;
    }
    return N_SHAPES;
}

int qppvm_create(const qppvm_desc* d, qppvm_handle** out)
{
    if (!d || !out) return fail(nullptr, QPPVM_ERR_ARG, "null argument");
    *out = nullptr;
    qppvm_layout L;
    if (qppvm_get_layout(d, &L)) return fail(nullptr, QPPVM_ERR_ARG, "invalid problem description");
    if (d->max_iter < 1 || d->n_reg_steps < 0 || !(d->eps_regularisation >= 0.0))
        return fail(nullptr, QPPVM_ERR_ARG, "invalid solver options");
    if (!(d->lambda_solver >= 0.0) || !(d->task_weight[0] >= 0.0) || !(d->task_weight[1] >= 0.0) || !(d->task_weight[2] >= 0.0))
        return fail(nullptr, QPPVM_ERR_ARG, "lambda_solver and the task weights must be positive (0 = default 1.0)");
    if (d->kind == QPPVM_KIND_TORQUE && (d->postural_actuated_only || (d->lambda_solver != 0.0 && d->lambda_solver != 1.0) ||
                                         (d->task_weight[0] != 0.0 && d->task_weight[0] != 1.0) || (d->task_weight[1] != 0.0 && d->task_weight[1] != 1.0) ||
                                         (d->task_weight[2] != 0.0 && d->task_weight[2] != 1.0)))
        return fail(nullptr, QPPVM_ERR_UNSUPPORTED, "lambda_solver / task weights / postural switch apply to the ForceAcc kind");
    const ShapeEntry* sh = nullptr;
    for (int i = 0; i < N_SHAPES; ++i)
        if (g_shapes[i].kind == d->kind && g_shapes[i].n_a == d->n_a && g_shapes[i].n_c == d->n_contacts && g_shapes[i].flags == d->flags)
            sh = &g_shapes[i];
    if (!sh) return fail(nullptr, QPPVM_ERR_UNSUPPORTED, "no sm_100a kernel instantiated for kind=%d n_a=%d contacts=%d flags=%d",
                         d->kind, d->n_a, d->n_contacts, d->flags);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= d->device)
        return fail(nullptr, QPPVM_ERR_NO_DEVICE, "no CUDA device %d (this library has no CPU path)", d->device);
    qppvm_handle* h = new (std::nothrow) qppvm_handle();
    if (!h) return fail(nullptr, QPPVM_ERR_ARG, "out of memory");
    memset(h, 0, sizeof(*h));
    h->desc = *d; h->L = L; h->shape = sh;
    // every failure below releases what has been created so far through qppvm_destroy (all members start out null)
#define CUC(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fail(nullptr, QPPVM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));        \
            qppvm_destroy(h);                                                                     \
            return QPPVM_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)
    DeviceGuard guard_(d->device);
    if (!guard_.ok) { fail(nullptr, QPPVM_ERR_CUDA, "cudaSetDevice(%d) failed", d->device); delete h; return QPPVM_ERR_CUDA; }
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, d->device));
    h->sm_count = prop.multiProcessorCount;
    // 64 threads per problem: one thread per variable / task column (n_x <= 64); 32-thread teams were measured
    // 28 % slower (two passes per loop and twice as many independent instruction streams per SM)
    h->team = 64;
    h->kernel = sh->kernel;
    CUC(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sh->slab_bytes));
    CUC(cudaFuncSetAttribute(h->kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int occ = 0;
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, h->kernel, h->team, (size_t)sh->slab_bytes));
    if (occ < 1) { fail(nullptr, QPPVM_ERR_CUDA, "kernel does not fit on an SM (%d B smem)", sh->slab_bytes); qppvm_destroy(h); return QPPVM_ERR_CUDA; }
    h->ctas_per_sm = occ;
    if (const char* e = getenv("QPPVM_CTAS_PER_SM")) { const int c = atoi(e); if (c >= 1 && c < occ) h->ctas_per_sm = c; }   // profiling aid
    if (sh->factor_kernel) {
        CUC(cudaFuncSetAttribute(sh->factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sh->factor_bytes));
        CUC(cudaFuncSetAttribute(sh->factor_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sh->factor_kernel, sh->factor_threads, (size_t)sh->factor_bytes));
        h->factor_ctas_per_sm = occ < 1 ? 1 : occ;
        const size_t cbytes = sizeof(double) * (size_t)L.rec_doubles;
        CUC(cudaFuncSetAttribute(sh->certify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cbytes));
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sh->certify_kernel, CERT_THREADS, cbytes));
        h->certify_ctas_per_sm = occ < 1 ? 1 : occ;
    }
    CUC(cudaMalloc(&h->counters, sizeof(unsigned long long) * N_SLOTS));
    // host path: records per pipelined chunk (H2D of chunk i+1 overlaps the solve of chunk i); QPPVM_CHUNK overrides
    if (const char* e = getenv("QPPVM_ROWWISE_EQUALITIES")) h->rowwise = atoi(e) != 0;
    h->chunk = 2048;      // (configs[2] end to end: 1024 -> 2.62 M, 2048 -> 2.80 M, 4096 -> 2.78 M solves/s; profiles/README.md)
    if (const char* e = getenv("QPPVM_CHUNK")) { const long c = atol(e); if (c >= 64 && c <= (1 << 20)) h->chunk = c; }
    h->chunk_states = 4 * h->chunk;
    if (const char* e = getenv("QPPVM_CHUNK_STATES")) { const long c = atol(e); if (c >= h->chunk && c <= (1 << 20)) h->chunk_states = c; }
    for (int i = 0; i < HOST_STREAMS; ++i) {
        CUC(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
        CUC(cudaMalloc(&h->d_rec[i], sizeof(double) * L.rec_doubles * h->chunk_states));
        CUC(cudaMalloc(&h->d_out[i], (size_t)L.out_bytes * h->chunk_states));
    }
    CUC(cudaStreamCreateWithFlags(&h->one_stream, cudaStreamNonBlocking));
    CUC(cudaMalloc(&h->d_one_rec, sizeof(double) * L.rec_doubles));
    CUC(cudaMalloc(&h->d_one_out, L.out_bytes));
    CUC(cudaMallocHost(&h->h_one_rec, sizeof(double) * L.rec_doubles));
    CUC(cudaMallocHost(&h->h_one_out, L.out_bytes));
    CUC(cudaEventCreateWithFlags(&h->ev_dev, cudaEventDisableTiming));
    CUC(cudaMalloc(&h->d_one_warm, sizeof(uint32_t) * QPPVM_WARM_WORDS));
    CUC(cudaMemset(h->d_one_warm, 0, sizeof(uint32_t) * QPPVM_WARM_WORDS));
    h->resident = sh->factor_kernel != nullptr;                 // latency mode through resident kernels (split shapes)
    if (const char* e = getenv("QPPVM_RESIDENT")) h->resident = h->resident && atoi(e) != 0;
    h->idle_us = 20000;
    if (const char* e = getenv("QPPVM_TICK_IDLE_US")) { const long v = atol(e); if (v >= 100 && v <= 10000000) h->idle_us = (uint32_t)v; }
    if (h->resident) {
        CUC(cudaHostAlloc(&h->tick_host, 64, cudaHostAllocPortable));
        memset(h->tick_host, 0, 64);
        CUC(cudaMalloc(&h->tick_dev, 128));
        CUC(cudaMemset(h->tick_dev, 0, 128));
        for (int i = 0; i < 3; ++i) CUC(cudaStreamCreateWithFlags(&h->tick_streams[i], cudaStreamNonBlocking));
    }
    if (sh->factor_kernel) {
        // Prepare workspaces, one per launch slot, sized for the largest pass that slot ever runs (nothing is allocated
        // on the solve path): the host-path chunks, WS_CHUNK problems for the caller's stream, one problem for the
        // latency slot.  The bulk copies of the solve kernel also move the unused Q1 / RN entries: defined contents.
        for (int i = 0; i < N_SLOTS; ++i) {
            int64_t ws_chunk = WS_CHUNK;                       // QPPVM_WS_CHUNK: measurement aid (profiles/README.md: pass-size sweep)
            if (const char* e = getenv("QPPVM_WS_CHUNK")) { const long c = atol(e); if (c >= 256 && c <= (1 << 20)) ws_chunk = c; }
            const int64_t capn = i < HOST_STREAMS ? h->chunk_states : (i == HOST_STREAMS ? ws_chunk : 8);
            const size_t bytes = sizeof(double) * (size_t)sh->ws_doubles * capn;
            CUC(cudaMalloc(&h->ws[i], bytes));
            CUC(cudaMemset(h->ws[i], 0, bytes));
            h->ws_cap[i] = capn;
        }
    }
#undef CUC
    *out = h;
    return QPPVM_OK;
}

int qppvm_destroy(qppvm_handle* h)
{
    if (!h) return QPPVM_ERR_ARG;
    DeviceGuard guard_(h->desc.device);
    stop_servers(h);
    for (int i = 0; i < 3; ++i) if (h->tick_streams[i]) cudaStreamDestroy(h->tick_streams[i]);
    if (h->tick_host) cudaFreeHost(h->tick_host);
    cudaFree(h->tick_dev); cudaFree(h->d_one_warm);
    for (int i = 0; i < HOST_STREAMS; ++i) {
        if (h->streams[i]) { cudaStreamSynchronize(h->streams[i]); cudaStreamDestroy(h->streams[i]); }
        cudaFree(h->d_rec[i]); cudaFree(h->d_out[i]);
    }
    if (h->one_stream) cudaStreamDestroy(h->one_stream);
    for (int i = 0; i < HOST_STREAMS; ++i) cudaFree(h->d_state[i]);
    cudaFree(h->rob_blob); cudaFree(h->d_roll); cudaFree(h->d_roll_warm);
    if (h->ev_fork) { cudaEventDestroy(h->ev_fork); for (int i = 0; i < HOST_STREAMS; ++i) cudaEventDestroy(h->ev_join[i]); }
    cudaFree(h->d_one_rec); cudaFree(h->d_one_out);
    cudaFreeHost(h->h_one_rec); cudaFreeHost(h->h_one_out);
    cudaFree(h->counters);
    for (int i = 0; i < N_SLOTS; ++i) cudaFree(h->ws[i]);
    if (h->ev_dev) cudaEventDestroy(h->ev_dev);
    if (h->tev) { for (cudaEvent_t e : *h->tev) cudaEventDestroy(e); delete h->tev; delete h->tkind; }
    delete h;
    return QPPVM_OK;
}

const char* qppvm_last_error(const qppvm_handle* h) { return h ? h->err : g_create_error; }

int qppvm_solve_batch_diag(qppvm_handle* h, const double* rec, void* out, double* diag, int64_t batch, void* stream)
{
    return solve_on_caller_stream(h, rec, out, diag, nullptr, false, batch, (cudaStream_t)stream);
}

int qppvm_solve_batch(qppvm_handle* h, const double* rec, void* out, int64_t batch, void* stream)
{
    return solve_on_caller_stream(h, rec, out, nullptr, nullptr, false, batch, (cudaStream_t)stream);
}

int qppvm_solve_batch_warm(qppvm_handle* h, const double* rec, void* out, uint32_t* warm, int64_t batch, void* stream)
{
    return solve_on_caller_stream(h, rec, out, nullptr, warm, true, batch, (cudaStream_t)stream);
}

int qppvm_solve_batch_host(qppvm_handle* h, const double* rec, void* out, int64_t batch)
{
    const int rc = qppvm_solve_batch_host_async(h, rec, out, batch);
    return rc ? rc : qppvm_host_sync(h);
}

int qppvm_host_sync(qppvm_handle* h)
{
    if (!h) return QPPVM_ERR_ARG;
    ENTER(h);
    for (int i = 0; i < HOST_STREAMS; ++i) CU(h, cudaStreamSynchronize(h->streams[i]));
    return QPPVM_OK;
}

int qppvm_solve_batch_host_async(qppvm_handle* h, const double* rec, void* out, int64_t batch)
{
    if (!h) return QPPVM_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!rec || !out))) return fail(h, QPPVM_ERR_ARG, "bad batch arguments");
    ENTER(h);
    const size_t rb = sizeof(double) * h->L.rec_doubles, ob = (size_t)h->L.out_bytes;
    int s = 0;
    for (int64_t c0 = 0; c0 < batch; c0 += h->chunk, s = (s + 1) % HOST_STREAMS) {
        const int64_t n = batch - c0 < h->chunk ? batch - c0 : h->chunk;
        cudaStream_t st = h->streams[s];
        CU(h, cudaMemcpyAsync(h->d_rec[s], (const char*)rec + c0 * rb, n * rb, cudaMemcpyHostToDevice, st));
        int rc = launch(h, h->d_rec[s], h->d_out[s], nullptr, n, st, s);
        if (rc) return rc;
        CU(h, cudaMemcpyAsync((char*)out + c0 * ob, h->d_out[s], n * ob, cudaMemcpyDeviceToHost, st));
    }
    return QPPVM_OK;                     // per-stream staging buffers are reused in stream order: safe across calls
}

int qppvm_solve_one(qppvm_handle* h, const double* rec, void* out)
{
    if (!h || !rec || !out) return h ? fail(h, QPPVM_ERR_ARG, "null argument") : QPPVM_ERR_ARG;
    ENTER(h);
    const size_t rb = sizeof(double) * h->L.rec_doubles, ob = (size_t)h->L.out_bytes;
    memcpy(h->h_one_rec, rec, rb);
    if (h->resident) {
        // Latency mode: nothing is launched per tick.  The record goes into pinned memory, the tick number is bumped,
        // the resident prepare -> solve -> certify chain (struct Tick) picks it up and writes the output record and the
        // tick number back.  The solve starts from the working sets of the previous tick (qpOASES hot start,
        // ref:src/QPPVMPlugin.cpp:246); qppvm_reset_warm() forgets them.  The servers leave after idle_us without
        // a tick (so that they never block a device-wide synchronisation for long) and are started again on demand.
        if (!h->tick_running || __atomic_load_n(&h->tick_host[3], __ATOMIC_ACQUIRE) == 0u) {
            const int rc = start_servers(h);
            if (rc) return rc;
        }
        const uint32_t seq = ++h->tick_seq;
        __atomic_store_n(&h->tick_host[0], seq, __ATOMIC_RELEASE);
        for (unsigned long spins = 0;; ++spins) {
            if (__atomic_load_n(&h->tick_host[1], __ATOMIC_ACQUIRE) == seq) break;
            if ((spins & 1023) == 1023) {
                if (__atomic_load_n(&h->tick_host[3], __ATOMIC_ACQUIRE) == 0u &&
                    __atomic_load_n(&h->tick_host[1], __ATOMIC_ACQUIRE) != seq) {      // the chain left just before the tick
                    const int rc = start_servers(h);
                    if (rc) return rc;
                }
                if (cudaStreamQuery(h->tick_streams[1]) != cudaErrorNotReady && cudaPeekAtLastError() != cudaSuccess)
                    return fail(h, QPPVM_ERR_CUDA, "resident solve server failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
        }
        memcpy(out, h->h_one_out, ob);
        return QPPVM_OK;
    }
    // Shapes without the resident chain: one launch sequence per tick.  The record sits in pinned host memory that the
    // GPU addresses directly (UVA) and the outputs are stored straight back into pinned host memory.
    int rc = launch(h, h->h_one_rec, h->h_one_out, nullptr, 1, h->one_stream, HOST_STREAMS + 1, false, h->d_one_warm);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->one_stream));
    memcpy(out, h->h_one_out, ob);
    return QPPVM_OK;
}

int qppvm_tick_stamps(qppvm_handle* h, uint64_t* ns7)
{
    if (!h || !ns7) return QPPVM_ERR_ARG;
    if (!h->resident) return fail(h, QPPVM_ERR_UNSUPPORTED, "this shape has no resident latency chain");
    ENTER(h);
    CU(h, cudaMemcpyAsync(ns7, h->tick_dev + 8, 7 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->one_stream));
    CU(h, cudaStreamSynchronize(h->one_stream));
    return QPPVM_OK;
}

int qppvm_reset_warm(qppvm_handle* h)
{
    if (!h) return QPPVM_ERR_ARG;
    ENTER(h);
    // (no tick is in flight: the handle is used by one thread)  A blocking memset would wait for the resident servers.
    CU(h, cudaMemsetAsync(h->d_one_warm, 0, sizeof(uint32_t) * QPPVM_WARM_WORDS, h->one_stream));
    CU(h, cudaStreamSynchronize(h->one_stream));
    return QPPVM_OK;
}

static int state_layout(const qppvm_desc* d, RbdShape* sh)
{
    qppvm_layout L;
    if (!d || d->kind != QPPVM_KIND_FORCEACC || (d->flags & (QPPVM_FLAG_FULL_WRENCH | QPPVM_FLAG_COM_TASK)) || qppvm_get_layout(d, &L)) return -1;
    const int na = L.n_a, c = L.n_c;
    int o = 0;
    RbdShape s;
    memset(&s, 0, sizeof(s));
    s.n_a = na; s.n_v = L.n_v; s.n_c = c; s.flags = d->flags;
    s.off_jwaist = L.off_jwaist; s.off_jc = L.off_jc; s.off_M = L.off_M; s.off_h = L.off_h; s.off_jdqd = L.off_jdqd;
    s.off_rhs = L.off_rhs; s.off_taulim = L.off_taulim; s.off_cone = L.off_cone; s.off_fbox = L.off_fbox; s.rec_doubles = L.rec_doubles;
    s.s_q = o; o += na; s.s_qd = o; o += na; s.s_R0 = o; o += 9; s.s_p0 = o; o += 3; s.s_tw = o; o += 6; s.s_gains = o; o += 4;
    s.s_ori = o; o += 3; s.s_foot = o; o += 6 * c; s.s_mu = o; o += c; s.s_tscale = o; o += na;
    s.s_wpos = o; o += 3;
    s.state_doubles = o;
    if (sh) *sh = s;
    return o;
}

int qppvm_state_doubles(const qppvm_desc* d) { return state_layout(d, nullptr); }

int qppvm_set_robot(qppvm_handle* h, const qppvm_robot* r)
{
    if (!h || !r) return QPPVM_ERR_ARG;
    if (state_layout(&h->desc, &h->rsh) < 0) return fail(h, QPPVM_ERR_UNSUPPORTED, "the state front end covers the ForceAcc kind with 3-D contact forces and without the CoM force task");
    const int na = h->desc.n_a, nb = na + 1, nc = h->desc.n_contacts;
    if (r->n_a != na || nb > RBD_MAXB) return fail(h, QPPVM_ERR_ARG, "robot has %d joints, handle expects %d (max %d bodies)", r->n_a, na, RBD_MAXB);
    std::vector<int> depth(nb, 0);
    std::vector<unsigned long long> anc(nb, 0ull);
    int maxd = 0;
    for (int i = 1; i < nb; ++i) {
        const int pa = r->parent[i];
        if (pa < 0 || pa >= i) return fail(h, QPPVM_ERR_ARG, "bodies must be topologically ordered (parent[i] < i)");
        depth[i] = depth[pa] + 1; anc[i] = anc[pa] | (1ull << (i - 1));
        if (depth[i] > maxd) maxd = depth[i];
    }
    for (int i = 0; i < nc; ++i)
        if (r->contact_body[i] < 0 || r->contact_body[i] >= nb) return fail(h, QPPVM_ERR_ARG, "bad contact body");
    // one blob: doubles first (8-byte aligned), then the integer tables
    const size_t nd = (size_t)nb * 3 * 4 + nb + 2 * na;              // axis offset com inertia | mass | q_home tau_max
    const size_t bytes = nd * 8 + (size_t)nb * 8 + (size_t)(2 * nb + 4) * 4;
    std::vector<unsigned char> host(bytes, 0);
    double* hd = reinterpret_cast<double*>(host.data());
    size_t o = 0;
    auto putd = [&](const double* src, size_t n) { memcpy(hd + o, src, n * 8); const size_t at = o; o += n; return at; };
    const size_t o_axis = putd(r->axis, nb * 3), o_off = putd(r->offset, nb * 3), o_com = putd(r->com, nb * 3), o_in = putd(r->inertia, nb * 3);
    const size_t o_mass = putd(r->mass, nb), o_qh = putd(r->q_home, na), o_tm = putd(r->tau_max, na);
    unsigned long long* hanc = reinterpret_cast<unsigned long long*>(host.data() + nd * 8);
    memcpy(hanc, anc.data(), nb * 8);
    int* hi = reinterpret_cast<int*>(host.data() + nd * 8 + (size_t)nb * 8);
    memcpy(hi, r->parent, nb * 4); memcpy(hi + nb, depth.data(), nb * 4);
    for (int i = 0; i < 4; ++i) hi[2 * nb + i] = i < nc ? r->contact_body[i] : 0;
    ENTER(h);
    if (h->rob_blob) { cudaFree(h->rob_blob); h->rob_blob = nullptr; }
    CU(h, cudaMalloc(&h->rob_blob, bytes));
    CU(h, cudaMemcpy(h->rob_blob, host.data(), bytes, cudaMemcpyHostToDevice));
    const double* dd = reinterpret_cast<const double*>(h->rob_blob);
    const int* di = reinterpret_cast<const int*>((const unsigned char*)h->rob_blob + nd * 8 + (size_t)nb * 8);
    h->rob.n_a = na; h->rob.n_b = nb; h->rob.max_depth = maxd;
    h->rob.axis = dd + o_axis; h->rob.offset = dd + o_off; h->rob.com = dd + o_com; h->rob.inertia = dd + o_in;
    h->rob.mass = dd + o_mass; h->rob.q_home = dd + o_qh; h->rob.tau_max = dd + o_tm;
    h->rob.anc = reinterpret_cast<const unsigned long long*>((const unsigned char*)h->rob_blob + nd * 8);
    h->rob.parent = di; h->rob.depth = di + nb; h->rob.contact_body = di + 2 * nb;
    for (int i = 0; i < HOST_STREAMS; ++i)
        if (!h->d_state[i]) CU(h, cudaMalloc(&h->d_state[i], sizeof(double) * h->rsh.state_doubles * h->chunk_states));
    // launch geometry of the warp-per-state front end: RBD_WARPS states per CTA, per-body data in dynamic shared memory
    h->rbd_v1 = false;
    if (const char* e = getenv("QPPVM_RBD_V1")) h->rbd_v1 = atoi(e) != 0;
    h->rbd_smem = (int)(sizeof(double) * RBD_WARPS * rbd_warp_doubles(h->rsh));
    CU(h, cudaFuncSetAttribute(rbd_records_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->rbd_smem));
    CU(h, cudaFuncSetAttribute(rbd_records_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int rocc = 0;
    CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rocc, rbd_records_warp_kernel, 32 * RBD_WARPS, (size_t)h->rbd_smem));
    h->rbd_ctas_per_sm = rocc < 1 ? 1 : rocc;
    h->has_robot = true;
    return QPPVM_OK;
}

static int launch_rbd(qppvm_handle* h, const double* states, double* recs, int64_t batch, cudaStream_t st)
{
    if (batch <= 0) return QPPVM_OK;
    if (h->rbd_v1) {                                                    // QPPVM_RBD_V1=1: the thread-per-state kernel (A/B measurements)
        const int grid = (int)((batch + RBD_TEAM - 1) / RBD_TEAM);
        rbd_records_kernel<<<grid, RBD_TEAM, 0, st>>>(h->rob, h->rsh, states, recs, (long long)batch);
    } else {                                                            // a warp per state, per-body data in shared memory
        const long long need = (batch + RBD_WARPS - 1) / RBD_WARPS, cap = (long long)h->sm_count * h->rbd_ctas_per_sm;
        rbd_records_warp_kernel<<<(int)(need < cap ? need : cap), 32 * RBD_WARPS, h->rbd_smem, st>>>(h->rob, h->rsh, states, recs, (long long)batch);
    }
    CU(h, cudaGetLastError());
    h->launches += 1;
    return QPPVM_OK;
}

int qppvm_records_from_states(qppvm_handle* h, const double* states, double* recs, int64_t batch, void* stream)
{
    if (!h) return QPPVM_ERR_ARG;
    if (!h->has_robot) return fail(h, QPPVM_ERR_ARG, "qppvm_set_robot has not been called");
    if (batch < 0 || (batch > 0 && (!states || !recs))) return fail(h, QPPVM_ERR_ARG, "bad batch arguments");
    ENTER(h);
    return launch_rbd(h, states, recs, batch, (cudaStream_t)stream);
}

static int launch_integrate(qppvm_handle* h, double* states, const void* out, const double* recs, double dt, int64_t batch, cudaStream_t st)
{
    if (batch <= 0) return QPPVM_OK;
    const int threads = 128, grid = (int)((batch + threads - 1) / threads);
    integrate_states_kernel<<<grid, threads, 0, st>>>(h->rsh, states, (const double*)out, recs, h->L.out_bytes / 8,
                                                      h->L.n_x + h->L.n_a, dt, (long long)batch);
    CU(h, cudaGetLastError());
    h->launches += 1;
    return QPPVM_OK;
}

int qppvm_integrate_states_tracking(qppvm_handle* h, double* states, const void* out, const double* recs, double dt,
                                    int64_t batch, void* stream)
{
    if (!h) return QPPVM_ERR_ARG;
    if (!h->has_robot) return fail(h, QPPVM_ERR_ARG, "qppvm_set_robot has not been called");
    if (batch < 0 || (batch > 0 && (!states || !out)) || !(dt > 0.0)) return fail(h, QPPVM_ERR_ARG, "bad arguments");
    ENTER(h);
    return launch_integrate(h, states, out, recs, dt, batch, (cudaStream_t)stream);
}

int qppvm_integrate_states(qppvm_handle* h, double* states, const void* out, double dt, int64_t batch, void* stream)
{
    return qppvm_integrate_states_tracking(h, states, out, nullptr, dt, batch, stream);
}

int qppvm_rollout_states(qppvm_handle* h, double* states, void* out, int ticks, double dt, int64_t batch, void* stream)
{
    if (!h) return QPPVM_ERR_ARG;
    if (!h->has_robot) return fail(h, QPPVM_ERR_ARG, "qppvm_set_robot has not been called");
    if (batch < 0 || ticks < 0 || (batch > 0 && (!states || !out)) || !(dt > 0.0)) return fail(h, QPPVM_ERR_ARG, "bad arguments");
    if (batch == 0 || ticks == 0) return QPPVM_OK;
    ENTER(h);
    cudaStream_t st = (cudaStream_t)stream;
    // States are independent, so the batch is cut into lanes that run all their ticks back to back on the handle's
    // worker streams: the tail of one lane's solve overlaps the front end / solve of the others (a single stream
    // leaves the SMs idle in every kernel's tail: 4.0 M -> state-ticks/s measured in profiles/README.md).
    int64_t lane = ((batch + HOST_STREAMS - 1) / HOST_STREAMS + 63) / 64 * 64;
    if (lane < 512) lane = 512;
    if (lane > ROLL_LANE) lane = ROLL_LANE;
    if (lane > h->roll_cap) {
        CU(h, cudaDeviceSynchronize());
        cudaFree(h->d_roll); h->d_roll = nullptr; h->roll_cap = 0;
        CU(h, cudaMalloc(&h->d_roll, sizeof(double) * (size_t)h->L.rec_doubles * lane * HOST_STREAMS));
        h->roll_cap = lane;
    }
    if (!h->ev_fork) {
        CU(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        for (int i = 0; i < HOST_STREAMS; ++i) CU(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
    }
    // Every tick of a state starts from the working sets its previous tick ended with, which is what the reference gets
    // from keeping one QPOases_sot alive across control_loop calls (ref:src/ForceAcc.cpp:135-137,189); the first tick is cold.
    if (batch > h->roll_warm_cap) {
        CU(h, cudaDeviceSynchronize());
        cudaFree(h->d_roll_warm); h->d_roll_warm = nullptr; h->roll_warm_cap = 0;
        CU(h, cudaMalloc(&h->d_roll_warm, sizeof(uint32_t) * QPPVM_WARM_WORDS * (size_t)batch));
        h->roll_warm_cap = batch;
    }
    CU(h, cudaMemsetAsync(h->d_roll_warm, 0, sizeof(uint32_t) * QPPVM_WARM_WORDS * (size_t)batch, st));
    const size_t sd = h->rsh.state_doubles, ob = (size_t)h->L.out_bytes;
    CU(h, cudaEventRecord(h->ev_fork, st));
    int used = 0, w = 0;
    for (int64_t c0 = 0; c0 < batch; c0 += lane, w = (w + 1) % HOST_STREAMS) {
        const int64_t n = batch - c0 < lane ? batch - c0 : lane;
        cudaStream_t ws = h->streams[w];
        if (used < HOST_STREAMS && w == used) { CU(h, cudaStreamWaitEvent(ws, h->ev_fork, 0)); ++used; }
        double* s = states + c0 * sd;
        double* rec = h->d_roll + (size_t)w * h->roll_cap * h->L.rec_doubles;
        unsigned char* o = (unsigned char*)out + c0 * ob;
        for (int t = 0; t < ticks; ++t) {
            int rc = launch_rbd(h, s, rec, n, ws);
            if (rc) return rc;
            rc = launch(h, rec, o, nullptr, n, ws, w, true, h->d_roll_warm + c0 * QPPVM_WARM_WORDS);
            if (rc) return rc;
            rc = launch_integrate(h, s, o, rec, dt, n, ws);
            if (rc) return rc;
        }
    }
    for (int i = 0; i < used; ++i) {
        CU(h, cudaEventRecord(h->ev_join[i], h->streams[i]));
        CU(h, cudaStreamWaitEvent(st, h->ev_join[i], 0));
    }
    return QPPVM_OK;
}

int qppvm_solve_states_host(qppvm_handle* h, const double* states, void* out, int64_t batch)
{
    const int rc = qppvm_solve_states_host_async(h, states, out, batch);
    return rc ? rc : qppvm_host_sync(h);
}

int qppvm_solve_states_host_async(qppvm_handle* h, const double* states, void* out, int64_t batch)
{
    if (!h) return QPPVM_ERR_ARG;
    if (!h->has_robot) return fail(h, QPPVM_ERR_ARG, "qppvm_set_robot has not been called");
    if (batch < 0 || (batch > 0 && (!states || !out))) return fail(h, QPPVM_ERR_ARG, "bad batch arguments");
    ENTER(h);
    const size_t sb = sizeof(double) * h->rsh.state_doubles, ob = (size_t)h->L.out_bytes;
    int s = 0;
    // chunk: a quarter of the batch (so that the four streams overlap copy / front end / solve), between the record
    // path's chunk and four times that (large batches: fewer, fuller launches)
    int64_t cs = (batch / 4 + 255) / 256 * 256;
    if (cs < h->chunk) cs = h->chunk;
    if (cs > h->chunk_states) cs = h->chunk_states;
    for (int64_t c0 = 0; c0 < batch; c0 += cs, s = (s + 1) % HOST_STREAMS) {
        const int64_t n = batch - c0 < cs ? batch - c0 : cs;
        cudaStream_t st = h->streams[s];
        CU(h, cudaMemcpyAsync(h->d_state[s], (const char*)states + c0 * sb, n * sb, cudaMemcpyHostToDevice, st));
        int rc = launch_rbd(h, h->d_state[s], h->d_rec[s], n, st);
        if (rc) return rc;
        rc = launch(h, h->d_rec[s], h->d_out[s], nullptr, n, st, s);
        if (rc) return rc;
        CU(h, cudaMemcpyAsync((char*)out + c0 * ob, h->d_out[s], n * ob, cudaMemcpyDeviceToHost, st));
    }
    return QPPVM_OK;
}

int64_t qppvm_kernel_launches(const qppvm_handle* h) { return h ? h->launches : 0; }

int qppvm_kernel_timing(qppvm_handle* h, int enable, double* ms3, int64_t* launches3)
{
    if (!h) return QPPVM_ERR_ARG;
    ENTER(h);
    if (!h->tev) { h->tev = new std::vector<cudaEvent_t>(); h->tkind = new std::vector<int>(); }
    if (ms3 && launches3) {
        for (int k = 0; k < 3; ++k) { ms3[k] = 0.0; launches3[k] = 0; }
        for (size_t i = 0; i < h->tkind->size(); ++i) {
            cudaEvent_t e0 = (*h->tev)[2 * i], e1 = (*h->tev)[2 * i + 1];
            float ms = 0.f;
            if (cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) {
                ms3[(*h->tkind)[i]] += ms; launches3[(*h->tkind)[i]] += 1;
            }
        }
    }
    for (cudaEvent_t e : *h->tev) cudaEventDestroy(e);
    h->tev->clear(); h->tkind->clear();
    h->timing = enable != 0;
    return QPPVM_OK;
}

int qppvm_reserve_sms(qppvm_handle* h, int n_sms)
{
    if (!h) return QPPVM_ERR_ARG;
    if (n_sms < 0 || n_sms >= h->sm_count) return fail(h, QPPVM_ERR_ARG, "cannot reserve %d of %d SMs", n_sms, h->sm_count);
    h->reserve_sms = n_sms;
    return QPPVM_OK;
}

int qppvm_fp64_peak(qppvm_handle* h, double* tflops)
{
    if (!h || !tflops) return QPPVM_ERR_ARG;
    ENTER(h);
    const int blocks = h->sm_count * 8, threads = 256, iters = 1 << 16;
    double* buf = nullptr;
    CU(h, cudaMalloc(&buf, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CU(h, cudaEventCreate(&e0)); CU(h, cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads>>>(buf, 1024);           // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(h, cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(buf, iters);
        CU(h, cudaEventRecord(e1));
        CU(h, cudaEventSynchronize(e1));
        float ms = 0;
        CU(h, cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    h->launches += 4;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return QPPVM_OK;
}

}  // extern "C"
