// qppvm_multi.cu — the batch sharded over the GPUs of one box, single process, behind the C-ABI.
//
// The path shards trivially (every state's QP cascade is independent: ref:src/QPPVMPlugin.cpp:246,
// ref:src/ForceAcc.cpp:189 solve exactly one per tick), so there is no data-path collective: rank r of G owns the
// contiguous block [r B / G, (r + 1) B / G).  NCCL only moves data to and from a root GPU -- grouped ncclSend / ncclRecv
// over NVLink / NVSwitch -- and that movement is pipelined against the solves: the block of every non-root GPU is cut
// into chunks, and while chunk i is being solved chunk i + 1 is being scattered and chunk i - 1 gathered (separate
// communicators and streams for the two directions, triple-buffered staging, events for the hand-offs).  The root
// solves its own block in place, chunk by chunk, and keeps a few SMs free for NCCL's copy kernels.  With host buffers no GPU-to-GPU traffic is needed at all:
// every GPU pulls its own block over its own PCIe link.
//
// NCCL is loaded with dlopen at qppvm_multi_create (no link-time dependency: the single-GPU library loads without it,
// and inside a process that already carries torch's NCCL the same libnccl.so.2 is reused).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "../../include/qppvm_b200.h"

#ifndef QPPVM_MULTI_ROOT_SHARE_DEFAULT
#define QPPVM_MULTI_ROOT_SHARE_DEFAULT 1.0      // tuned on 8 GPUs: see profiles/README.md
#endif
namespace {

constexpr int MAX_DEV = 16;
constexpr int NBUF = 3;      // staging depth: the scatter runs up to two chunks ahead of the solve

struct NcclApi {
    void* lib;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};

char g_multi_error[512] = "";

}  // namespace

struct qppvm_multi {
    int n;
    int dev[MAX_DEV];
    qppvm_handle* h[MAX_DEV];
    qppvm_layout L;
    NcclApi nccl;
    ncclComm_t scatter[MAX_DEV], gather[MAX_DEV];   // two communicators: the two directions run concurrently
    bool have_comms;
    cudaStream_t s_scatter[MAX_DEV], s_solve[MAX_DEV], s_gather[MAX_DEV];
    double* rec[MAX_DEV][NBUF];                     // staging on the non-root GPUs, NBUF chunks deep
    double* st[MAX_DEV][NBUF];                      // states form: the scattered compact states (allocated by set_robot)
    double* rec0;                                   // states form: the root's records of one chunk
    int state_doubles;                              // 0 until qppvm_multi_set_robot
    qppvm_desc desc;
    unsigned char* out[MAX_DEV][NBUF];
    cudaEvent_t ev_recv[MAX_DEV][NBUF], ev_solved[MAX_DEV][NBUF], ev_sent[MAX_DEV][NBUF], ev_root;
    int64_t chunk;                                  // records per pipeline chunk and GPU
    double root_share;                              // device-root form: the root's block relative to an equal share (QPPVM_MULTI_ROOT_SHARE)
    int64_t nccl_calls;
    char err[512];
};

namespace {

int mfail(qppvm_multi* m, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(m ? m->err : g_multi_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define MCU(m, call)                                                                          \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return mfail(m, QPPVM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define MNC(m, call)                                                                          \
    do {                                                                                      \
        ncclResult_t r_ = (call);                                                             \
        if (r_ != ncclSuccess) return mfail(m, QPPVM_ERR_CUDA, "%s failed: %s", #call, (m)->nccl.GetErrorString(r_)); \
    } while (0)

bool load_nccl(NcclApi* a)
{
    memset(a, 0, sizeof(*a));
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        a->lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a->lib) break;
    }
    if (!a->lib) return false;
    a->CommInitAll = (decltype(a->CommInitAll))dlsym(a->lib, "ncclCommInitAll");
    a->CommDestroy = (decltype(a->CommDestroy))dlsym(a->lib, "ncclCommDestroy");
    a->Send = (decltype(a->Send))dlsym(a->lib, "ncclSend");
    a->Recv = (decltype(a->Recv))dlsym(a->lib, "ncclRecv");
    a->GroupStart = (decltype(a->GroupStart))dlsym(a->lib, "ncclGroupStart");
    a->GroupEnd = (decltype(a->GroupEnd))dlsym(a->lib, "ncclGroupEnd");
    a->GetErrorString = (decltype(a->GetErrorString))dlsym(a->lib, "ncclGetErrorString");
    return a->CommInitAll && a->CommDestroy && a->Send && a->Recv && a->GroupStart && a->GroupEnd && a->GetErrorString;
}

struct DevGuard {
    int prev = -1;
    DevGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// block of rank r: [lo, hi)
inline void block_of(int64_t batch, int n, int r, int64_t* lo, int64_t* hi)
{
    *lo = batch * r / n;
    *hi = batch * (r + 1) / n;
}
// The same with the root's block scaled by `share` (device-root form: the root GPU also feeds every other GPU, its copy
// kernels and HBM reads compete with its own solves, so an equal split makes it the straggler); the other GPUs split the
// rest evenly.  Problems are independent: the split does not change a single output bit.
inline void block_of_root(int64_t batch, int n, int r, double share, int64_t* lo, int64_t* hi)
{
    if (n == 1 || share == 1.0) { block_of(batch, n, r, lo, hi); return; }
    int64_t b0 = (int64_t)(share * (double)(batch / n));
    if (b0 < 0) b0 = 0;
    if (b0 > batch) b0 = batch;
    const int64_t rest = batch - b0;
    if (r == 0) { *lo = 0; *hi = b0; return; }
    *lo = b0 + rest * (r - 1) / (n - 1);
    *hi = b0 + rest * r / (n - 1);
}

}  // namespace

extern "C" {

const char* qppvm_multi_last_error(const qppvm_multi* m) { return m ? m->err : g_multi_error; }

int qppvm_multi_destroy(qppvm_multi* m)
{
    if (!m) return QPPVM_ERR_ARG;
    DevGuard guard;
    for (int r = 0; r < m->n; ++r) {
        cudaSetDevice(m->dev[r]);
        if (m->s_scatter[r]) { cudaStreamSynchronize(m->s_scatter[r]); cudaStreamSynchronize(m->s_solve[r]); cudaStreamSynchronize(m->s_gather[r]); }
        if (m->have_comms) { m->nccl.CommDestroy(m->scatter[r]); m->nccl.CommDestroy(m->gather[r]); }
        for (int b = 0; b < NBUF; ++b) {
            cudaFree(m->rec[r][b]); cudaFree(m->out[r][b]); cudaFree(m->st[r][b]);
            if (m->ev_recv[r][b]) cudaEventDestroy(m->ev_recv[r][b]);
            if (m->ev_solved[r][b]) cudaEventDestroy(m->ev_solved[r][b]);
            if (m->ev_sent[r][b]) cudaEventDestroy(m->ev_sent[r][b]);
        }
        if (m->s_scatter[r]) { cudaStreamDestroy(m->s_scatter[r]); cudaStreamDestroy(m->s_solve[r]); cudaStreamDestroy(m->s_gather[r]); }
        if (m->h[r]) qppvm_destroy(m->h[r]);
    }
    if (m->ev_root) { cudaSetDevice(m->dev[0]); cudaEventDestroy(m->ev_root); }
    if (m->rec0) { cudaSetDevice(m->dev[0]); cudaFree(m->rec0); }
    delete m;
    return QPPVM_OK;
}

int qppvm_multi_create(const qppvm_desc* desc, const int32_t* devices, int n_devices, qppvm_multi** out)
{
    if (!desc || !out || n_devices < 1 || n_devices > MAX_DEV) return mfail(nullptr, QPPVM_ERR_ARG, "bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return mfail(nullptr, QPPVM_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
    qppvm_multi* m = new (std::nothrow) qppvm_multi();
    if (!m) return mfail(nullptr, QPPVM_ERR_ARG, "out of memory");
    memset(m, 0, sizeof(*m));
    m->n = n_devices;
    m->desc = *desc;
    DevGuard guard;
    for (int r = 0; r < n_devices; ++r) {
        m->dev[r] = devices ? devices[r] : r;
        if (m->dev[r] < 0 || m->dev[r] >= ndev) { mfail(nullptr, QPPVM_ERR_NO_DEVICE, "no CUDA device %d", m->dev[r]); qppvm_multi_destroy(m); return QPPVM_ERR_NO_DEVICE; }
        for (int q = 0; q < r; ++q)
            if (m->dev[q] == m->dev[r]) { mfail(nullptr, QPPVM_ERR_ARG, "device %d listed twice", m->dev[r]); qppvm_multi_destroy(m); return QPPVM_ERR_ARG; }
    }
    if (qppvm_get_layout(desc, &m->L)) { mfail(nullptr, QPPVM_ERR_ARG, "invalid problem description"); qppvm_multi_destroy(m); return QPPVM_ERR_ARG; }
    // records per pipeline chunk: large chunks make efficient solve launches (pass-size sweep in profiles/README.md), small
    // ones shorten the fill and drain of the scatter pipeline.  Measured on 8 GPUs, where the root's NVLink egress is the
    // bound (~330 GB/s through grouped ncclSend, 18.6 GB per 2^20-record batch): 16 384 -> 51-54 ms, 32 768 -> 54-57 ms,
    // 65 536 -> 62 ms per batch; on 2 GPUs (solve-bound) 32 768 wins.
    m->chunk = n_devices >= 8 ? 16384 : 32768;
    if (const char* e = getenv("QPPVM_MULTI_CHUNK")) { const long c = atol(e); if (c >= 256 && c <= (1 << 20)) m->chunk = c; }
    m->root_share = QPPVM_MULTI_ROOT_SHARE_DEFAULT;
    if (const char* e = getenv("QPPVM_MULTI_ROOT_SHARE")) { const double v = atof(e); if (v >= 0.0 && v <= 1.0) m->root_share = v; }
#define MCC(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            mfail(nullptr, QPPVM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));             \
            qppvm_multi_destroy(m);                                                                     \
            return QPPVM_ERR_CUDA;                                                                      \
        }                                                                                               \
    } while (0)
    for (int r = 0; r < n_devices; ++r) {
        qppvm_desc d = *desc;
        d.device = m->dev[r];
        const int rc = qppvm_create(&d, &m->h[r]);
        if (rc) { mfail(nullptr, rc, "device %d: %s", m->dev[r], qppvm_last_error(nullptr)); qppvm_multi_destroy(m); return rc; }
        MCC(cudaSetDevice(m->dev[r]));
        // the copy kernels of NCCL must get onto the SMs between the solve launches: highest priority
        int prio_lo = 0, prio_hi = 0;
        MCC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        MCC(cudaStreamCreateWithPriority(&m->s_scatter[r], cudaStreamNonBlocking, prio_hi));
        MCC(cudaStreamCreateWithPriority(&m->s_solve[r], cudaStreamNonBlocking, prio_lo));
        MCC(cudaStreamCreateWithPriority(&m->s_gather[r], cudaStreamNonBlocking, prio_hi));
        for (int b = 0; b < NBUF; ++b) {
            MCC(cudaEventCreateWithFlags(&m->ev_recv[r][b], cudaEventDisableTiming));
            MCC(cudaEventCreateWithFlags(&m->ev_solved[r][b], cudaEventDisableTiming));
            MCC(cudaEventCreateWithFlags(&m->ev_sent[r][b], cudaEventDisableTiming));
            if (r > 0) {
                MCC(cudaMalloc(&m->rec[r][b], sizeof(double) * (size_t)m->L.rec_doubles * m->chunk));
                MCC(cudaMalloc(&m->out[r][b], (size_t)m->L.out_bytes * m->chunk));
            }
        }
    }
    MCC(cudaSetDevice(m->dev[0]));
    MCC(cudaEventCreateWithFlags(&m->ev_root, cudaEventDisableTiming));
#undef MCC
    if (n_devices > 1) {
        if (!load_nccl(&m->nccl)) { mfail(nullptr, QPPVM_ERR_UNSUPPORTED, "libnccl.so.2 not found (needed for more than one GPU): %s", dlerror()); qppvm_multi_destroy(m); return QPPVM_ERR_UNSUPPORTED; }
        ncclResult_t r1 = m->nccl.CommInitAll(m->scatter, n_devices, m->dev);
        ncclResult_t r2 = r1 == ncclSuccess ? m->nccl.CommInitAll(m->gather, n_devices, m->dev) : r1;
        if (r1 != ncclSuccess || r2 != ncclSuccess) {
            mfail(nullptr, QPPVM_ERR_CUDA, "ncclCommInitAll failed: %s", m->nccl.GetErrorString(r1 != ncclSuccess ? r1 : r2));
            if (r1 == ncclSuccess) for (int r = 0; r < n_devices; ++r) m->nccl.CommDestroy(m->scatter[r]);
            qppvm_multi_destroy(m);
            return QPPVM_ERR_CUDA;
        }
        m->have_comms = true;
        // Optionally keep SMs of the root free for NCCL's copy kernels (measured on 2 GPUs: not needed once both ends of
        // every transfer are gated on the same readiness event, 0.93 vs 0.92 of two independent GPUs)
        int rs = 0;
        if (const char* e = getenv("QPPVM_MULTI_RESERVE_SMS")) rs = atoi(e);
        if (rs > 0) qppvm_reserve_sms(m->h[0], rs);
    }
    *out = m;
    return QPPVM_OK;
}

int qppvm_multi_devices(const qppvm_multi* m) { return m ? m->n : 0; }
int64_t qppvm_multi_nccl_calls(const qppvm_multi* m) { return m ? m->nccl_calls : 0; }
int64_t qppvm_multi_kernel_launches(const qppvm_multi* m)
{
    int64_t s = 0;
    if (m) for (int r = 0; r < m->n; ++r) s += qppvm_kernel_launches(m->h[r]);
    return s;
}
int qppvm_multi_set_robot(qppvm_multi* m, const qppvm_robot* robot)
{
    if (!m || !robot) return QPPVM_ERR_ARG;
    for (int r = 0; r < m->n; ++r) {
        const int rc = qppvm_set_robot(m->h[r], robot);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[r], qppvm_last_error(m->h[r]));
    }
    // staging of the device-root states form (qppvm_multi_solve_states): the scattered states on every other GPU, the
    // root's records of one chunk -- allocated here, once, so that nothing is allocated on the solve path
    if (!m->state_doubles) {
        const int sd = qppvm_state_doubles(&m->desc);
        if (sd <= 0) return mfail(m, QPPVM_ERR_UNSUPPORTED, "the state front end covers the ForceAcc kind only");
        DevGuard guard;
        for (int r = 1; r < m->n; ++r) {
            MCU(m, cudaSetDevice(m->dev[r]));
            for (int b = 0; b < NBUF; ++b) MCU(m, cudaMalloc(&m->st[r][b], sizeof(double) * (size_t)sd * m->chunk));
        }
        MCU(m, cudaSetDevice(m->dev[0]));
        MCU(m, cudaMalloc(&m->rec0, sizeof(double) * (size_t)m->L.rec_doubles * m->chunk));
        m->state_doubles = sd;
    }
    return QPPVM_OK;
}

// Inputs and outputs live on the root GPU (devices[0]); synchronous: returns when `out_root` is complete.
// states == false: `in_root` holds records.  states == true: compact states (qppvm_state_doubles each); they are what
// travels (1 KB instead of 11-18 KB per problem: the root's NVLink egress stops being the bound), and
// every GPU runs the rigid-body front end on its own chunk before the solve.
static int multi_root(qppvm_multi* m, const double* in_root, void* out_root, int64_t batch, bool states)
{
    if (!m) return QPPVM_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!in_root || !out_root))) return mfail(m, QPPVM_ERR_ARG, "bad batch arguments");
    if (states && !m->state_doubles) return mfail(m, QPPVM_ERR_ARG, "qppvm_multi_set_robot has not been called");
    if (batch == 0) return QPPVM_OK;
    DevGuard guard;
    const size_t rd = states ? (size_t)m->state_doubles : (size_t)m->L.rec_doubles, ob = (size_t)m->L.out_bytes;
    const double* const rec_root = in_root;
    const int n = m->n;
    // the caller's work on the root device (default stream) comes first
    MCU(m, cudaSetDevice(m->dev[0]));
    MCU(m, cudaEventRecord(m->ev_root, 0));
    MCU(m, cudaStreamWaitEvent(m->s_scatter[0], m->ev_root, 0));
    MCU(m, cudaStreamWaitEvent(m->s_solve[0], m->ev_root, 0));
    MCU(m, cudaStreamWaitEvent(m->s_gather[0], m->ev_root, 0));
    int64_t lo0, hi0;
    block_of_root(batch, n, 0, m->root_share, &lo0, &hi0);
    // one chunk on one GPU: [front end ->] solve, in stream order
    auto solve_chunk = [&](int r, const double* in, double* recs, void* out, int64_t cn) -> int {
        if (states) {
            const int rc = qppvm_records_from_states(m->h[r], in, recs, cn, m->s_solve[r]);
            if (rc) return rc;
        }
        return qppvm_solve_batch(m->h[r], states ? recs : in, out, cn, m->s_solve[r]);
    };
    if (n == 1) {
        for (int64_t c0 = 0; c0 < batch; c0 += states ? m->chunk : batch) {
            const int64_t cn = states ? (batch - c0 < m->chunk ? batch - c0 : m->chunk) : batch;
            const int rc = solve_chunk(0, rec_root + c0 * rd, m->rec0, (unsigned char*)out_root + c0 * ob, cn);
            if (rc) return mfail(m, rc, "device %d: %s", m->dev[0], qppvm_last_error(m->h[0]));
        }
    }
    if (n > 1) {
        int64_t maxblk = hi0 - lo0;
        for (int r = 1; r < n; ++r) { int64_t lo, hi; block_of_root(batch, n, r, m->root_share, &lo, &hi); if (hi - lo > maxblk) maxblk = hi - lo; }
        const int64_t nchunks = (maxblk + m->chunk - 1) / m->chunk;
        // software pipeline over chunk index c: scatter(c), solve(c), gather(c) are enqueued in that order, each on its own
        // stream per GPU; the hardware overlaps scatter(c + 1) and gather(c - 1) with solve(c)
        for (int64_t c = 0; c < nchunks; ++c) {
            const int b = (int)(c % NBUF);
            // ---- scatter chunk c: root -> every other GPU (one group)
            MNC(m, m->nccl.GroupStart());
            for (int r = 1; r < n; ++r) {
                int64_t lo, hi;
                block_of_root(batch, n, r, m->root_share, &lo, &hi);
                const int64_t c0 = lo + c * m->chunk;
                if (c0 >= hi) continue;
                const int64_t cn = hi - c0 < m->chunk ? hi - c0 : m->chunk;
                MCU(m, cudaSetDevice(m->dev[r]));
                if (c >= NBUF) {
                    // staging buffer b is free again.  BOTH ends wait for it: a send kernel launched early would spin on the
                    // root's SMs until the matching receive is posted.
                    MCU(m, cudaStreamWaitEvent(m->s_scatter[r], m->ev_solved[r][b], 0));
                    MCU(m, cudaSetDevice(m->dev[0]));
                    MCU(m, cudaStreamWaitEvent(m->s_scatter[0], m->ev_solved[r][b], 0));
                    MCU(m, cudaSetDevice(m->dev[r]));
                }
                MNC(m, m->nccl.Recv(states ? m->st[r][b] : m->rec[r][b], (size_t)cn * rd, ncclDouble, 0, m->scatter[r], m->s_scatter[r]));
                MCU(m, cudaSetDevice(m->dev[0]));
                MNC(m, m->nccl.Send(rec_root + c0 * rd, (size_t)cn * rd, ncclDouble, r, m->scatter[0], m->s_scatter[0]));
                m->nccl_calls += 2;
            }
            MNC(m, m->nccl.GroupEnd());
            // ---- solve chunk c: the root in place (its launches are chunked too: a persistent kernel over the whole
            // block would keep NCCL's copy kernels off the root's SMs until it ends), every other GPU from its staging
            {
                const int64_t c0 = lo0 + c * m->chunk;
                if (c0 < hi0) {
                    const int64_t cn = hi0 - c0 < m->chunk ? hi0 - c0 : m->chunk;
                    const int rc = solve_chunk(0, rec_root + c0 * rd, m->rec0, (unsigned char*)out_root + c0 * ob, cn);
                    if (rc) return mfail(m, rc, "device %d: %s", m->dev[0], qppvm_last_error(m->h[0]));
                }
            }
            for (int r = 1; r < n; ++r) {
                int64_t lo, hi;
                block_of_root(batch, n, r, m->root_share, &lo, &hi);
                const int64_t c0 = lo + c * m->chunk;
                if (c0 >= hi) continue;
                const int64_t cn = hi - c0 < m->chunk ? hi - c0 : m->chunk;
                MCU(m, cudaSetDevice(m->dev[r]));
                MCU(m, cudaEventRecord(m->ev_recv[r][b], m->s_scatter[r]));
                MCU(m, cudaStreamWaitEvent(m->s_solve[r], m->ev_recv[r][b], 0));
                if (c >= NBUF) MCU(m, cudaStreamWaitEvent(m->s_solve[r], m->ev_sent[r][b], 0));       // output buffer b has been gathered
                const int rc = solve_chunk(r, states ? m->st[r][b] : m->rec[r][b], m->rec[r][b], m->out[r][b], cn);
                if (rc) return mfail(m, rc, "device %d: %s", m->dev[r], qppvm_last_error(m->h[r]));
                MCU(m, cudaEventRecord(m->ev_solved[r][b], m->s_solve[r]));
                MCU(m, cudaStreamWaitEvent(m->s_gather[r], m->ev_solved[r][b], 0));
                // (the root's receive kernel must not be resident, spinning, while this GPU is still solving)
                MCU(m, cudaSetDevice(m->dev[0]));
                MCU(m, cudaStreamWaitEvent(m->s_gather[0], m->ev_solved[r][b], 0));
            }
            // ---- gather chunk c: outputs straight into their place in the root's output block (one group)
            MNC(m, m->nccl.GroupStart());
            for (int r = 1; r < n; ++r) {
                int64_t lo, hi;
                block_of_root(batch, n, r, m->root_share, &lo, &hi);
                const int64_t c0 = lo + c * m->chunk;
                if (c0 >= hi) continue;
                const int64_t cn = hi - c0 < m->chunk ? hi - c0 : m->chunk;
                MCU(m, cudaSetDevice(m->dev[r]));
                MNC(m, m->nccl.Send(m->out[r][b], (size_t)cn * ob, ncclChar, 0, m->gather[r], m->s_gather[r]));
                MCU(m, cudaSetDevice(m->dev[0]));
                MNC(m, m->nccl.Recv((unsigned char*)out_root + c0 * ob, (size_t)cn * ob, ncclChar, r, m->gather[0], m->s_gather[0]));
                m->nccl_calls += 2;
            }
            MNC(m, m->nccl.GroupEnd());
            for (int r = 1; r < n; ++r) {
                int64_t lo, hi;
                block_of_root(batch, n, r, m->root_share, &lo, &hi);
                if (lo + c * m->chunk >= hi) continue;
                MCU(m, cudaSetDevice(m->dev[r]));
                MCU(m, cudaEventRecord(m->ev_sent[r][b], m->s_gather[r]));
            }
        }
    }
    for (int r = 0; r < n; ++r) {
        MCU(m, cudaSetDevice(m->dev[r]));
        MCU(m, cudaStreamSynchronize(m->s_scatter[r]));
        MCU(m, cudaStreamSynchronize(m->s_solve[r]));
        MCU(m, cudaStreamSynchronize(m->s_gather[r]));
    }
    return QPPVM_OK;
}

int qppvm_multi_solve_batch(qppvm_multi* m, const double* rec_root, void* out_root, int64_t batch)
{
    return multi_root(m, rec_root, out_root, batch, false);
}

int qppvm_multi_solve_states(qppvm_multi* m, const double* states_root, void* out_root, int64_t batch)
{
    return multi_root(m, states_root, out_root, batch, true);
}

// Host buffers (pinned for full speed): every GPU moves its own block over its own PCIe link; no GPU-to-GPU traffic.
static int multi_host(qppvm_multi* m, const void* in_host, void* out_host, int64_t batch, bool states)
{
    if (!m) return QPPVM_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!in_host || !out_host))) return mfail(m, QPPVM_ERR_ARG, "bad batch arguments");
    if (batch == 0) return QPPVM_OK;
    qppvm_desc d0;
    memset(&d0, 0, sizeof(d0));
    size_t in_stride = sizeof(double) * (size_t)m->L.rec_doubles;
    if (states) {
        // state stride from the first handle's description (qppvm_state_doubles is pure host arithmetic)
        d0.kind = QPPVM_KIND_FORCEACC; d0.n_a = m->L.n_a; d0.n_contacts = m->L.n_c;
        d0.flags = (m->L.row_cone >= 0 ? QPPVM_FLAG_FRICTION_CONES : 0) | (m->L.row_tau >= 0 ? QPPVM_FLAG_TORQUE_LIMITS : 0);
        const int sd = qppvm_state_doubles(&d0);
        if (sd <= 0) return mfail(m, QPPVM_ERR_UNSUPPORTED, "the state front end covers the ForceAcc kind only");
        in_stride = sizeof(double) * (size_t)sd;
    }
    const size_t ob = (size_t)m->L.out_bytes;
    for (int r = 0; r < m->n; ++r) {
        int64_t lo, hi;
        block_of(batch, m->n, r, &lo, &hi);
        const char* in = (const char*)in_host + lo * in_stride;
        char* o = (char*)out_host + lo * ob;
        const int rc = states ? qppvm_solve_states_host_async(m->h[r], (const double*)in, o, hi - lo)
                              : qppvm_solve_batch_host_async(m->h[r], (const double*)in, o, hi - lo);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[r], qppvm_last_error(m->h[r]));
    }
    for (int r = 0; r < m->n; ++r) {
        const int rc = qppvm_host_sync(m->h[r]);
        if (rc) return mfail(m, rc, "device %d: %s", m->dev[r], qppvm_last_error(m->h[r]));
    }
    return QPPVM_OK;
}

int qppvm_multi_solve_batch_host(qppvm_multi* m, const double* records_host, void* out_host, int64_t batch)
{
    return multi_host(m, records_host, out_host, batch, false);
}

int qppvm_multi_solve_states_host(qppvm_multi* m, const double* states_host, void* out_host, int64_t batch)
{
    return multi_host(m, states_host, out_host, batch, true);
}

}  // extern "C"
