// qcm_b200.cu -- sm_100a execution layer behind include/qcm_b200.h.
//
// What runs here replaces the dense inner loops of QCMaquis's contraction engine
// (dmrg/mp_tensors/contractions/{abelian,non-abelian,common}): the block dgemm calls of
// block_matrix_algorithms.h:48-162 / non-abelian/gemm.hpp:48-204, the W-application axpy panels of
// alps_detail.hpp:189-224 / micro_kernels.hpp:19-198, the pairing reshapes of reshapes.h:177-335 and the
// critical-section reduction over the MPO bond index (abelian/site_hamil.hpp:84-86).  The host flattens
// those loops into task arrays once per (site, direction); the kernels below execute them:
//
//   k_copy_panels   pairing reshapes as panel copies                                                   (this file)
//   k_gemm_ws       grouped, variable-size FP64 GEMM on DMMA, persistent and warp-specialised             (gemm_ws.cu)
//   k_wstream       the W application for destination panels with 2..4 source panels (FMA streaming)     (this file)
//   k_wgemm_ws      the W application for high fan-in panels as gathered dense products on DMMA           (wgemm_ws.cu)
//                   Destination panels with a single source are never formed: the closing GEMM reads the source.
//   k_vec_*         solver-side BLAS-1 on device-resident vectors
//
// There is no CPU fallback: every entry point fails with a status when the device is not usable.
#include "qcm_dev.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

// --------------------------------------------------------------------------------------------------------
// error handling
static thread_local std::string g_err;
static int fail(std::string const& s) { g_err = s; return 1; }
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define CHECK_INIT() do { if (!G.ready) return fail("qcm_init has not been called (or failed): no usable CUDA device"); } while (0)

struct qcm_array_s { double* p; int64_t n; bool resident = true; cudaEvent_t ready = nullptr; bool pending = false; };

// device-side task records ---------------------------------------------------------------------------------
struct DCopy { long long src_off, dst_off; int src_buf, dst_buf, rows, cols, lds, ldd; };

// --------------------------------------------------------------------------------------------------------
// kernels
__global__ void k_copy_panels(const DCopy* __restrict__ tasks, const __grid_constant__ BufTable bufs)
{
    DCopy t = tasks[blockIdx.x];
    const double* __restrict__ s = bufs.p[t.src_buf] + t.src_off;
    double* __restrict__ d = bufs.p[t.dst_buf] + t.dst_off;
    int n = t.rows * t.cols;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int r = i % t.rows, c = i / t.rows;
        d[r + (long long)c * t.ldd] = s[r + (long long)c * t.lds];
    }
}

// W application, low fan-in (at most 4 sources, at most 4 destinations with identical source sets): a streaming
// kernel, HBM bound by construction.  One CTA = WS_TILE consecutive panel elements of one group; every thread keeps
// WS_PT independent loads per source in flight.
constexpr int WS_THREADS = 128, WS_PT = 8, WS_TILE = WS_THREADS * WS_PT;
__global__ void __launch_bounds__(WS_THREADS)
k_wstream(const DWWork* __restrict__ works, const DWGroup* __restrict__ groups, const DWSrc* __restrict__ srcs,
          const DWDst* __restrict__ dsts, const double* __restrict__ coefs, const __grid_constant__ BufTable bufs)
{
    const DWWork w = works[blockIdx.x];
    const DWGroup g = groups[w.group];
    const int n = g.rows * g.cols;
    const double* __restrict__ sp[4]; int sl[4]; double cf[4][4];
    double* __restrict__ dp[4]; int dl[4];
    bool flat = true;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        sp[u] = nullptr; sl[u] = g.rows;
        if (u < g.n_src) { const DWSrc q = srcs[g.src_begin + u]; sp[u] = bufs.p[q.buf] + q.off; sl[u] = q.lds; flat = flat && q.lds == g.rows; }
#pragma unroll
        for (int d = 0; d < 4; ++d) cf[u][d] = (u < g.n_src && d < g.n_dst) ? coefs[g.coef_begin + u * 4 + d] : 0.;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        dp[d] = nullptr; dl[d] = g.rows;
        if (d < g.n_dst) { const DWDst q = dsts[g.dst_begin + d]; dp[d] = bufs.p[q.buf] + q.off; dl[d] = q.ldd; flat = flat && q.ldd == g.rows; }
    }
    double x[WS_PT][4];
    int rr[WS_PT], cc[WS_PT];
#pragma unroll
    for (int i = 0; i < WS_PT; ++i) {
        const int e = w.e0 + threadIdx.x + i * WS_THREADS;
        if (flat) { rr[i] = e; cc[i] = 0; } else { cc[i] = e / g.rows; rr[i] = e - cc[i] * g.rows; }
#pragma unroll
        for (int u = 0; u < 4; ++u) x[i][u] = (e < n && u < g.n_src) ? sp[u][rr[i] + (long long)cc[i] * sl[u]] : 0.;
    }
#pragma unroll
    for (int i = 0; i < WS_PT; ++i) {
        const int e = w.e0 + threadIdx.x + i * WS_THREADS;
        if (e >= n) continue;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d >= g.n_dst) continue;
            double a = x[i][0] * cf[0][d];
#pragma unroll
            for (int u = 1; u < 4; ++u) a = fma(x[i][u], cf[u][d], a);
            dp[d][rr[i] + (long long)cc[i] * dl[d]] = a;
        }
    }
}

// solver-side BLAS-1 --------------------------------------------------------------------------------------
// Dot products are reduced in a FIXED order (per-thread strided sums -> warp shuffles -> one partial per block -> one block
// sums the partials): the result depends on (n, grid) only, never on scheduling, so every rank of a sharded run computes
// bit-identical solver scalars from bit-identical vectors and all ranks take the same convergence decisions.
constexpr int DOT_THREADS = 256, DOT_MAX_BLOCKS = 1024;
__global__ void __launch_bounds__(DOT_THREADS) k_vec_dot_partial(const double* __restrict__ x, const double* __restrict__ y, long long n, double* __restrict__ partial)
{
    double s = 0.;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fma(x[i], y[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
// out[q] = sum of partial[q * stride .. q * stride + n_partial) for q = blockIdx.x
__global__ void __launch_bounds__(DOT_THREADS) k_vec_dot_final(const double* __restrict__ partial, int n_partial, int stride, double* __restrict__ out)
{
    const double* p = partial + (size_t)blockIdx.x * stride;
    double s = 0.;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) s += p[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[blockIdx.x] = s;
    }
}
__global__ void k_vec_axpy(double a, const double* __restrict__ x, double* __restrict__ y, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fma(a, x[i], y[i]);
}
__global__ void k_vec_scal(double a, double* __restrict__ x, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= a;
}

// peak probes ---------------------------------------------------------------------------------------------
__global__ void k_peak_fma(double* out, int iters)
{
    double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}
__global__ void k_peak_dmma(double* out, int iters)
{
    double c[8][2], a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma8x8x4(c[i][0], c[i][1], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
__global__ void k_copy_stream(const double2* __restrict__ s, double2* __restrict__ d, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}

// --------------------------------------------------------------------------------------------------------
// library state
struct GemmLaunch { int variant; int64_t work_begin, n_works; };   // works of one tile variant are contiguous
struct GemmGroup
{
    DWork* d_works = nullptr; DSeg* d_segs = nullptr;
    std::vector<GemmLaunch> launches;
    int64_t n_works = 0;
};
struct AxpyGroup
{
    DWWork* d_works = nullptr; DWGroup* d_groups = nullptr; DWSrc* d_srcs = nullptr; DWDst* d_dsts = nullptr; double* d_coefs = nullptr;
    int64_t begin[5] = {0, 0, 0, 0, 0}, count[5] = {0, 0, 0, 0, 0};   // work ranges: [0] stream, [1..4] DMMA product with ng = 8, 16, 32, 64
};
struct WaveDev { GemmGroup t, c; AxpyGroup w; int64_t y_elems = 0, t_elems = 0, x_chunk = 0; bool x_zero = false; };

struct qcm_plan_s
{
    int kind = 0;
    int rank = 0, world = 1;          // sharding the plan was built for (checked against the communicator at execution)
    DCopy* d_copies = nullptr; int64_t n_copies = 0;
    GemmGroup p;
    std::vector<WaveDev> waves;
    int64_t elems[QCM_BUF_COUNT];
    double flops = 0; int64_t bytes = 0;
    int64_t n_launches = 0;
    std::vector<void*> allocs;
    std::vector<size_t> alloc_caps;       // capacity of allocs[i] when it is a recyclable task buffer (flush_uploads)
    // task arrays are staged in the library's pinned buffer while the plan is built and go to the device in ONE allocation and one copy
    int64_t task_bytes = 0;
    struct PendingPtr { void** where; size_t offset; };
    std::vector<PendingPtr> pending;
};

typedef int (*nccl_get_uid_t)(void*);
typedef int (*nccl_init_rank_t)(void**, int, char[128], int);   // ncclUniqueId is passed BY VALUE (128-byte struct)
struct NcclUid { char internal[128]; };
typedef int (*nccl_get_uid_fn)(NcclUid*);
typedef int (*nccl_init_rank_fn)(void**, int, NcclUid, int);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_reducescatter_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_destroy_fn)(void*);
typedef const char* (*nccl_errstr_fn)(int);

static struct Global
{
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    static constexpr int kAux = 3;
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr};   // tile variants of one GEMM group run side by side
    cudaEvent_t fork_ev = nullptr, join_ev[kAux] = {nullptr, nullptr, nullptr};
    int64_t launches = 0;
    double* ws[QCM_BUF_COUNT] = {nullptr};
    int64_t ws_elems[QCM_BUF_COUNT] = {0};
    double* scratch = nullptr;          // small device scalar area
    double* dot_partial = nullptr;      // per-block partial sums of qcm_vec_dots + the results
    double* xacc = nullptr; int64_t xacc_elems = 0;   // accumulator of the exchange region (time-sliced shards)
    static constexpr int kLcSlots = 64;
    char* lc_args = nullptr; unsigned lc_next = 0;   // argument slots of qcm_vec_lincomb
    bool timing = false;
    double last_ms[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8];
    // nccl (resolved at run time so that single-GPU use has no NCCL dependency)
    void* nccl_lib = nullptr; void* comm = nullptr; int rank = 0, world = 1;
    nccl_get_uid_fn f_uid = nullptr; nccl_init_rank_fn f_init = nullptr; nccl_allreduce_fn f_ar = nullptr; nccl_reducescatter_fn f_rs = nullptr;
    cudaStream_t comm_stream = nullptr; cudaEvent_t x_ready = nullptr, x_done = nullptr;   // exchange of partial W sums, overlapped with the local waves
    cudaStream_t copy_stream = nullptr; cudaEvent_t spill_ev = nullptr;                    // spill tier (qcm_array_evict / prefetch)
    nccl_destroy_fn f_destroy = nullptr; nccl_errstr_fn f_err = nullptr;
} G;

static int gemm_set_attributes();
static int wgemm_set_attributes();
static void release_task_bufs();      // recycled task buffers of destroyed plans (plan construction, below)
static int ensure_ws(int slot, int64_t n)
{
    if (n <= G.ws_elems[slot]) return 0;
    if (G.ws[slot]) { CU(cudaStreamSynchronize(G.stream)); CU(cudaFree(G.ws[slot])); G.ws[slot] = nullptr; G.ws_elems[slot] = 0; }
    int64_t want = n + n / 8 + 1024;
    cudaError_t e = cudaMalloc((void**)&G.ws[slot], (size_t)want * sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = n;
        cudaStreamSynchronize(G.stream);
        release_task_bufs();
        cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, G.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        e = cudaMalloc((void**)&G.ws[slot], (size_t)want * sizeof(double));
        if (e != cudaSuccess) return fail("workspace allocation of " + std::to_string(n * 8) + " bytes failed: " + cudaGetErrorString(e));
    }
    G.ws_elems[slot] = want;
    return 0;
}

extern "C" int qcm_init(int device)
{
    if (G.ready && G.device == device) return 0;
    if (G.ready) return fail("qcm_init: already bound to device " + std::to_string(G.device));
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail("qcm_init: device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(std::string("device ") + prop.name + " is not sm_100-class; this library carries sm_100a code only");
    G.sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
    {
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;        // freed arrays stay in the pool for the next site
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto& ev : G.ev) CU(cudaEventCreate(&ev));
    for (int i = 0; i < Global::kAux; ++i) { CU(cudaStreamCreateWithFlags(&G.aux[i], cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&G.join_ev[i], cudaEventDisableTiming)); }
    CU(cudaEventCreateWithFlags(&G.fork_ev, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&G.comm_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&G.copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&G.spill_ev, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&G.x_ready, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&G.x_done, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&G.scratch, 4096));
    if (gemm_set_attributes() || wgemm_set_attributes()) return 1;
    G.device = device;
    G.ready = true;
    return 0;
}

extern "C" int qcm_finalize(void)
{
    if (!G.ready) return 0;
    cudaStreamSynchronize(G.stream);
    release_task_bufs();
    for (int i = 0; i < QCM_BUF_COUNT; ++i) if (G.ws[i]) { cudaFree(G.ws[i]); G.ws[i] = nullptr; G.ws_elems[i] = 0; }
    if (G.scratch) cudaFree(G.scratch);
    G.scratch = nullptr;
    if (G.dot_partial) cudaFree(G.dot_partial);
    G.dot_partial = nullptr;
    if (G.xacc) cudaFree(G.xacc);
    G.xacc = nullptr; G.xacc_elems = 0;
    if (G.lc_args) cudaFree(G.lc_args);
    G.lc_args = nullptr;
    for (auto& ev : G.ev) cudaEventDestroy(ev);
    for (int i = 0; i < Global::kAux; ++i) { cudaStreamDestroy(G.aux[i]); cudaEventDestroy(G.join_ev[i]); }
    cudaEventDestroy(G.fork_ev);
    cudaEventDestroy(G.x_ready); cudaEventDestroy(G.x_done); cudaStreamDestroy(G.comm_stream);
    cudaStreamSynchronize(G.copy_stream); cudaStreamDestroy(G.copy_stream); cudaEventDestroy(G.spill_ev);
    cudaStreamDestroy(G.stream);
    G.stream = nullptr; G.ready = false; G.device = -1;
    return 0;
}

extern "C" const char* qcm_last_error(void) { return g_err.c_str(); }
// for the host-side translation units of the library (plan_capi.cpp)
extern "C" int qcm_internal_fail(const char* msg) { return fail(msg ? msg : "error"); }
extern "C" void qcm_internal_forget_plan(qcm_plan_t p);
extern "C" int qcm_device_count(int* n)
{
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; return fail(cudaGetErrorString(e)); }
    return 0;
}
extern "C" int qcm_device_name(char* buf, int len)
{
    CHECK_INIT();
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, G.device));
    snprintf(buf, len, "%s", prop.name);
    return 0;
}
extern "C" int qcm_sync(void) { CHECK_INIT(); CU(cudaStreamSynchronize(G.stream)); return 0; }
extern "C" void* qcm_stream(void) { return (void*)G.stream; }
extern "C" int64_t qcm_launch_count(void) { return G.launches; }

// ---- arrays ---------------------------------------------------------------------------------------------
extern "C" int qcm_array_alloc(int64_t n, qcm_array_t* out)
{
    CHECK_INIT();
    if (n < 0) return fail("qcm_array_alloc: negative size");
    qcm_array_s* a = new qcm_array_s(); a->p = nullptr; a->n = n;
    if (n > 0) {
        // stream-ordered allocation from the device's default pool (its release threshold is raised in qcm_init): boundaries
        // and solver vectors come and go at every site of a sweep without a device-wide synchronisation
        const size_t bytes = (size_t)n * sizeof(double);
        cudaError_t e = cudaMallocAsync((void**)&a->p, bytes, G.stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaStreamSynchronize(G.stream);
            release_task_bufs();
            cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, G.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
            e = cudaMallocAsync((void**)&a->p, (size_t)n * sizeof(double), G.stream);
        }
        if (e != cudaSuccess) { cudaGetLastError(); delete a; return fail(std::string("qcm_array_alloc: ") + cudaGetErrorString(e)); }
    }
    *out = a;
    return 0;
}
extern "C" int qcm_array_free(qcm_array_t a)
{
    if (!a) return 0;
    if (a->p) { if (G.ready) cudaFreeAsync(a->p, G.stream); else cudaFree(a->p); }
    if (a->ready) cudaEventDestroy(a->ready);
    delete a;
    return 0;
}
// ---- spill tier ---------------------------------------------------------------------------------------------
extern "C" int qcm_pinned_alloc(int64_t n, double** out)
{
    CHECK_INIT();
    if (n < 0 || !out) return fail("qcm_pinned_alloc: bad argument");
    *out = nullptr;
    if (n == 0) return 0;
    CU(cudaHostAlloc((void**)out, (size_t)n * sizeof(double), cudaHostAllocDefault));
    return 0;
}
extern "C" int qcm_pinned_free(double* p) { if (p) cudaFreeHost(p); return 0; }
extern "C" int qcm_array_resident(qcm_array_t a, int* resident) { if (!a || !resident) return fail("null argument"); *resident = a->resident ? 1 : 0; return 0; }
extern "C" int qcm_array_evict(qcm_array_t a, double* host)
{
    CHECK_INIT();
    if (!a) return fail("qcm_array_evict: null array");
    if (!a->resident) return 0;
    if (a->n > 0) {
        if (!host) return fail("qcm_array_evict: null host buffer");
        // after everything queued on the compute stream (the array's producers and readers), on the copy stream
        CU(cudaEventRecord(G.spill_ev, G.stream));
        CU(cudaStreamWaitEvent(G.copy_stream, G.spill_ev, 0));
        CU(cudaMemcpyAsync(host, a->p, (size_t)a->n * sizeof(double), cudaMemcpyDeviceToHost, G.copy_stream));
        CU(cudaFreeAsync(a->p, G.copy_stream));
        a->p = nullptr;
    }
    a->resident = false; a->pending = false;
    return 0;
}
extern "C" int qcm_array_prefetch(qcm_array_t a, const double* host)
{
    CHECK_INIT();
    if (!a) return fail("qcm_array_prefetch: null array");
    if (a->resident) return 0;
    if (a->n > 0) {
        if (!host) return fail("qcm_array_prefetch: null host buffer");
        CU(cudaMallocAsync((void**)&a->p, (size_t)a->n * sizeof(double), G.copy_stream));
        CU(cudaMemcpyAsync(a->p, host, (size_t)a->n * sizeof(double), cudaMemcpyHostToDevice, G.copy_stream));
        if (!a->ready) CU(cudaEventCreateWithFlags(&a->ready, cudaEventDisableTiming));
        CU(cudaEventRecord(a->ready, G.copy_stream));
        a->pending = true;
    }
    a->resident = true;
    return 0;
}
extern "C" int qcm_array_size(qcm_array_t a, int64_t* n) { if (!a) return fail("null array"); *n = a->n; return 0; }
extern "C" void* qcm_array_devptr(qcm_array_t a) { return a ? (void*)a->p : nullptr; }
extern "C" int qcm_array_upload(qcm_array_t a, int64_t off, const double* host, int64_t n)
{
    CHECK_INIT();
    if (!a || off < 0 || n < 0 || off + n > a->n) return fail("qcm_array_upload: range outside the array");
    if (n == 0) return 0;
    if (!a->resident) return fail("qcm_array_upload: the array has been evicted");
    if (a->pending) { CU(cudaStreamWaitEvent(G.stream, a->ready, 0)); a->pending = false; }
    CU(cudaMemcpyAsync(a->p + off, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}
extern "C" int qcm_array_download(qcm_array_t a, int64_t off, double* host, int64_t n)
{
    CHECK_INIT();
    if (!a || off < 0 || n < 0 || off + n > a->n) return fail("qcm_array_download: range outside the array");
    if (n == 0) return 0;
    if (!a->resident) return fail("qcm_array_download: the array has been evicted");
    if (a->pending) { CU(cudaStreamWaitEvent(G.stream, a->ready, 0)); a->pending = false; }
    CU(cudaMemcpyAsync(host, a->p + off, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}
extern "C" int qcm_array_zero(qcm_array_t a)
{
    CHECK_INIT();
    if (!a) return fail("null array");
    if (!a->resident) return fail("qcm_array_zero: the array has been evicted");
    if (a->pending) { CU(cudaStreamWaitEvent(G.stream, a->ready, 0)); a->pending = false; }
    if (a->n) CU(cudaMemsetAsync(a->p, 0, (size_t)a->n * sizeof(double), G.stream));
    return 0;
}

// ---- plan construction ----------------------------------------------------------------------------------
constexpr int KC = 16;   // K chunk of the GEMM kernels (gemm_ws.cu)
static int gemm_set_attributes()
{
    const char* e = gemm_ws_init(G.sm_count);
    if (e) return fail(std::string("gemm_ws_init: ") + e);
    return 0;
}

// Cost model (SM cycles) of one tile of em x en with `chunks` K chunks.  The FP64 pipe issues one DMMA (8x8x4) per 4
// cycles and SM: 16 cycles per 8x8 fragment and chunk of 16; a chunk also has to be staged through L2.  For the strip
// classes the per-chunk cost is the one MEASURED on B200 at cfg3 (profiles/r01c_counters_cfg3.txt: launch time * SMs /
// chunks): thin strips are bound by operand traffic, not by the pipe.  Every tile pays its epilogue.
static double tile_cost(int em, int en, int chunks)
{
    const int em8 = (em + 7) & ~7, en8 = (en + 7) & ~7;
    const double frags = (double)(em8 / 8) * (en8 / 8);
    auto cls = [](int x) { return x > 64 ? 128 : x > 32 ? 64 : x > 16 ? 32 : x > 8 ? 16 : 8; };
    const int a = std::min(cls(em8), cls(en8)), b = std::max(cls(em8), cls(en8));
    double per_chunk;
    if (b == 128) per_chunk = a == 128 ? 4520. : a == 64 ? 2500. : a == 32 ? 1500. : a == 16 ? 1150. : 1650.;
    else per_chunk = std::max(16.0 * frags, 8.0 * (em8 + en8)) + 300.0;
    return chunks * per_chunk + 4.0 * frags + 300.0;
}

// A dimension of an output block is cut into strips of 128 plus strips of {64, 32, 16, 8} (or one more 128) that
// cover the remainder at the least modelled cost: the kernels always compute whole warp tiles, so a strip class that
// fits the remainder tightly wastes no FP64 issue slots, while many thin strips waste operand bandwidth.
static void cut_dimension(int len, int chunks, std::vector<std::pair<int, int>>& strips /* (offset, class) */)
{
    static const int cls[5] = {8, 16, 32, 64, 128};
    strips.clear();
    int off = 0;
    while (len - off >= 128) { strips.push_back(std::make_pair(off, 128)); off += 128; }
    int r8 = (len - off + 7) / 8;      // fragments left (0..15)
    if (r8 == 0) return;
    // dynamic programme over "fragments still to cover"
    double best[17]; int pick[17];
    best[0] = 0; pick[0] = -1;
    for (int r = 1; r <= 16; ++r) {
        best[r] = 1e300; pick[r] = 4;
        for (int c = 0; c < 5; ++c) {
            const int f = cls[c] / 8;
            const double cost = tile_cost(cls[c], 128, chunks) + best[std::max(0, r - f)];
            if (cost < best[r]) { best[r] = cost; pick[r] = c; }
        }
    }
    std::vector<int> sel;
    for (int r = r8; r > 0; r = std::max(0, r - cls[pick[r]] / 8)) sel.push_back(cls[pick[r]]);
    std::sort(sel.begin(), sel.end(), [](int a, int b) { return a > b; });
    for (int c : sel) { strips.push_back(std::make_pair(off, c)); off += c; }
}
// tile variant (gemm_ws.cu table) for a (row strip class, column strip class) pair
static int variant_for(int hr, int hc)
{
    if (hr == 128 && hc == 128) return 0;
    if (hc == 128) return hr == 64 ? 1 : hr == 32 ? 3 : hr == 16 ? 5 : 7;
    if (hr == 128) return hc == 64 ? 2 : hc == 32 ? 4 : hc == 16 ? 6 : 8;
    const int mx = std::max(hr, hc), mn = std::min(hr, hc);
    if (mx == 64) return mn == 64 ? 9 : mn == 32 ? (hr == 64 ? 12 : 13) : (hr == 64 ? 14 : 15);
    return mx == 32 ? 10 : 11;
}

// Task arrays go to the device in ONE allocation and one copy.  They are gathered in a pinned host buffer owned by the
// library (grown on demand, reused by every plan): the copy then runs at PCIe speed instead of through the driver's
// pageable staging, and the large arrays (the W coefficient tables, 0.4 GB for a cfg3 centre site) are copied exactly once
// on the host, by several threads.
static struct PinnedStage { char* p = nullptr; size_t cap = 0, used = 0; } g_pin;
static std::vector<std::pair<char*, size_t>> g_task_bufs;      // device buffers of destroyed plans, kept for the next plans (at most kTaskBufsKept)
constexpr size_t kTaskBufsKept = 3;
static void release_task_bufs()
{
    for (auto& b : g_task_bufs) cudaFree(b.first);
    g_task_bufs.clear();
}
static int pin_reserve(size_t need)
{
    if (need <= g_pin.cap) return 0;
    size_t cap = std::max(need + need / 4, (size_t)1 << 24);
    char* q = nullptr;
    cudaError_t e = cudaHostAlloc((void**)&q, cap, cudaHostAllocDefault);
    if (e != cudaSuccess) { cap = need; e = cudaHostAlloc((void**)&q, cap, cudaHostAllocDefault); }
    if (e != cudaSuccess) return fail(std::string("pinned staging buffer of ") + std::to_string(need) + " bytes: " + cudaGetErrorString(e));
    if (g_pin.used) memcpy(q, g_pin.p, g_pin.used);
    if (g_pin.p) cudaFreeHost(g_pin.p);
    g_pin.p = q; g_pin.cap = cap;
    return 0;
}
static void big_memcpy(char* dst, const char* src, size_t n)
{
    const size_t kMin = (size_t)32 << 20;
    if (n < 2 * kMin) { memcpy(dst, src, n); return; }
    const int nt = (int)std::min<size_t>(8, n / kMin);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        size_t b = n * t / nt, e = n * (t + 1) / nt;
        th.emplace_back([=]() { memcpy(dst + b, src + b, e - b); });
    }
    for (auto& x : th) x.join();
}
static int dev_upload_raw(qcm_plan_s* P, const void* src, size_t bytes, void** d)
{
    *d = nullptr;
    if (bytes == 0) return 0;
    size_t off = (g_pin.used + 255) & ~(size_t)255;
    if (pin_reserve(off + bytes)) return 1;
    big_memcpy(g_pin.p + off, (const char*)src, bytes);
    g_pin.used = off + bytes;
    P->pending.push_back(qcm_plan_s::PendingPtr{d, off});
    return 0;
}
template <class T> static int dev_upload(qcm_plan_s* P, std::vector<T> const& h, T** d) { return dev_upload_raw(P, h.data(), h.size() * sizeof(T), (void**)d); }
// one cudaMalloc + one copy for all task arrays of the plan; patches the device pointers recorded by dev_upload
static int flush_uploads(qcm_plan_s* P)
{
    if (g_pin.used == 0) return 0;
    char* base = nullptr;
    // A sweep creates and destroys two plans per site.  Their task buffers are recycled (g_task_bufs: the buffer of a destroyed plan
    // serves the next plan that fits; all use is ordered on the library's stream) -- asking the device pool for half a gigabyte of a
    // new size at every site costs tens of milliseconds each time.
    size_t cap = 0;
    {
        size_t best = g_task_bufs.size();
        for (size_t i = 0; i < g_task_bufs.size(); ++i)
            if (g_task_bufs[i].second >= g_pin.used && g_task_bufs[i].second <= 4 * g_pin.used + ((size_t)16 << 20) &&
                (best == g_task_bufs.size() || g_task_bufs[i].second < g_task_bufs[best].second)) best = i;
        if (best != g_task_bufs.size()) { base = g_task_bufs[best].first; cap = g_task_bufs[best].second; g_task_bufs.erase(g_task_bufs.begin() + (long)best); }
    }
    if (!base) {
        cap = g_pin.used + g_pin.used / 4;
        cudaError_t em = cudaMallocAsync((void**)&base, cap, G.stream);
        if (em != cudaSuccess) {
            cudaGetLastError();
            cudaStreamSynchronize(G.stream);
            release_task_bufs();
            cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, G.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
            cap = g_pin.used;
            CU(cudaMallocAsync((void**)&base, cap, G.stream));
        }
    }
    P->allocs.push_back(base);
    P->alloc_caps.push_back(cap);
    P->task_bytes = (int64_t)g_pin.used;
    CU(cudaMemcpyAsync(base, g_pin.p, g_pin.used, cudaMemcpyHostToDevice, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    for (auto const& q : P->pending) *q.where = base + q.offset;
    g_pin.used = 0;
    P->pending.clear();
    return 0;
}

// Host threads of plan creation.  qcm_plan_create sits between two device phases of a sweep (the solver of one site and the
// next), so its loops over 10^6 outputs / segments run on several threads; ranges are contiguous and their results are
// concatenated in range order, so the task arrays do not depend on the number of threads.
static int plan_threads()
{
    static const int n = []() {
        if (const char* e = getenv("QCM_PLAN_THREADS")) return std::max(1, atoi(e));
        // one process per GPU: a rank takes its share of the host cores, the share the launcher gave its OpenMP runtime
        if (const char* e = getenv("OMP_NUM_THREADS")) { int n = atoi(e); if (n >= 1) return std::min(16, n); }
        unsigned hc = std::thread::hardware_concurrency();
        return (int)std::min<unsigned>(16u, std::max<unsigned>(1u, hc));
    }();
    return n;
}
template <class F> static void parallel_ranges(int64_t n, int64_t min_per_thread, F f /* (int t, int64_t begin, int64_t end) */, int* n_ranges = nullptr)
{
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(plan_threads(), n / std::max<int64_t>(1, min_per_thread)));
    if (n_ranges) *n_ranges = nt;
    if (nt == 1) { f(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back([=]() { f(t, n * t / nt, n * (t + 1) / nt); });
    f(0, (int64_t)0, n / nt);
    for (auto& x : th) x.join();
}

// output blocks -> tile work items; long segment lists are split (split-K over the MPO bond index) and
// combined with FP64 atomics
static int build_gemm_group(qcm_plan_s* P, GemmGroup& g, const qcm_gemm_out* outs, int64_t n_outs, const qcm_gemm_seg* segs, int64_t n_segs,
                            int base_mode /*0 store, 1 add*/)
{
    std::vector<DSeg> hs((size_t)n_segs);
    parallel_ranges(n_segs, 1 << 16, [&](int, int64_t b, int64_t e) {
        for (int64_t i = b; i < e; ++i) {
            qcm_gemm_seg const& s = segs[i];
            hs[i] = DSeg{s.A.off, s.B.off, s.A.buf, s.B.buf, s.lda, s.ldb, s.m, s.n, s.k, s.ta, s.tb, 0, s.alpha};
        }
    });
    const int kNumVariants = gemm_ws_num_variants();
    typedef std::vector<std::vector<std::pair<double, DWork>>> PerVariant;
    const int max_chunks = 96;       // K chunks (of KC) per work item before the segment list is split
    std::vector<PerVariant> parts((size_t)plan_threads(), PerVariant((size_t)kNumVariants));
    int n_parts = 1;
    parallel_ranges(n_outs, 1 << 12, [&](int t, int64_t ob, int64_t oe) {
        PerVariant& per_variant = parts[(size_t)t];
        std::vector<std::pair<int, int>> chunks, rows, cols;
        for (int64_t o = ob; o < oe; ++o) {
            qcm_gemm_out const& out = outs[o];
            if (out.m <= 0 || out.n <= 0) continue;
            chunks.clear();
            {
                int csum = 0, cb = out.seg_begin;
                for (int s = out.seg_begin; s < out.seg_end; ++s) {
                    csum += (segs[s].k + KC - 1) / KC;
                    // only accumulating outputs are split: a store-mode output (step-1 products) runs its whole K list in one work item
                    if (base_mode == 1 && csum >= max_chunks && s + 1 < out.seg_end) { chunks.push_back(std::make_pair(cb, s + 1)); cb = s + 1; csum = 0; }
                }
                chunks.push_back(std::make_pair(cb, out.seg_end));
            }
            int total_chunks = 0;
            for (int s = out.seg_begin; s < out.seg_end; ++s) total_chunks += (segs[s].k + KC - 1) / KC;
            const int avg_chunks = std::max(1, total_chunks / (int)chunks.size());
            int mode = chunks.size() > 1 ? 2 : base_mode;
            cut_dimension(out.m, avg_chunks, rows);
            cut_dimension(out.n, avg_chunks, cols);
            for (auto const& ch : chunks) {
                int nch = 0;
                for (int s = ch.first; s < ch.second; ++s) nch += (segs[s].k + KC - 1) / KC;
                for (auto const& cs : cols)
                    for (auto const& rs : rows) {
                        const int v = variant_for(rs.second, cs.second);
                        const GemmWsVariant var = gemm_ws_variant(v);
                        // a variant tile may be larger than the strip pair (corner pieces): it is clipped by the block edge
                        per_variant[v].push_back(std::make_pair(tile_cost(std::min(var.tm, out.m - rs.first), std::min(var.tn, out.n - cs.first), nch),
                                                                DWork{out.C.off, out.C.buf, out.ldc, rs.first, cs.first, out.m, out.n, ch.first, ch.second, mode, 0}));
                    }
            }
        }
    }, &n_parts);
    // ranges in order: the same sequence per variant as a single pass over the outputs
    PerVariant per_variant((size_t)kNumVariants);
    for (int v = 0; v < kNumVariants; ++v) {
        size_t tot = 0;
        for (int t = 0; t < n_parts; ++t) tot += parts[(size_t)t][(size_t)v].size();
        if (n_parts == 1) { per_variant[(size_t)v].swap(parts[0][(size_t)v]); continue; }
        per_variant[(size_t)v].reserve(tot);
        for (int t = 0; t < n_parts; ++t) per_variant[(size_t)v].insert(per_variant[(size_t)v].end(), parts[(size_t)t][(size_t)v].begin(), parts[(size_t)t][(size_t)v].end());
    }
    parts.clear();
    // the persistent CTAs take work items round-robin: heaviest first, so that every CTA gets a similar mix and the
    // launch ends on light items; the launches of a group are ordered by total cost
    std::vector<DWork> hw;
    std::vector<std::pair<double, int>> order;
    for (int v = 0; v < kNumVariants; ++v) {
        if (per_variant[v].empty()) continue;
        double tot = 0;
        for (auto const& x : per_variant[v]) tot += x.first;
        order.push_back(std::make_pair(-tot, v));
    }
    std::sort(order.begin(), order.end());
    for (auto const& ov : order) {
        const int v = ov.second;
        // heaviest first -- to the resolution that matters for load balance: a stable counting sort over 1/8-octave cost classes
        // (a comparison sort of the 10^6 work items of a cfg3 closing group took longer than planning the boundary step)
        {
            auto const& items = per_variant[v];
            constexpr int kBins = 8 * 64;
            auto bin_of = [](double c) { int e = 0; double m = std::frexp(c > 1. ? c : 1., &e); int b = e * 8 + (int)((m - 0.5) * 16.); return kBins - 1 - std::min(kBins - 1, std::max(0, b)); };
            std::vector<uint32_t> start(kBins + 1, 0);
            for (auto const& x : items) start[(size_t)bin_of(x.first) + 1]++;
            for (int b = 0; b < kBins; ++b) start[(size_t)b + 1] += start[(size_t)b];
            const size_t base = hw.size();
            hw.resize(base + items.size());
            for (auto const& x : items) hw[base + start[(size_t)bin_of(x.first)]++] = x.second;
            g.launches.push_back(GemmLaunch{v, (int64_t)base, (int64_t)items.size()});
        }
        if (getenv("QCM_DEBUG")) {      // useful vs issued FLOPs of this launch (issued = whole warp tiles, K padded to 4)
            const GemmWsVariant var = gemm_ws_variant(v);
            double useful = 0, chunks_n = 0;
            for (auto const& x : per_variant[v]) {
                DWork const& w = x.second;
                const int em = std::min(var.tm, w.m - w.m0), en = std::min(var.tn, w.n - w.n0);
                for (int sgi = w.seg_begin; sgi < w.seg_end; ++sgi) {
                    const int sm_ = std::min(em, segs[sgi].m - w.m0), sn_ = std::min(en, segs[sgi].n - w.n0);
                    if (sm_ > 0 && sn_ > 0) { useful += 2.0 * sm_ * sn_ * segs[sgi].k; chunks_n += (segs[sgi].k + KC - 1) / KC; }
                }
            }
            fprintf(stderr, "gemm launch: variant %2d (%3d x %3d) works %7zu chunks %9.0f useful %.4e flops, modelled cost %.3e cycles\n", v, var.tm, var.tn,
                    per_variant[v].size(), chunks_n, useful, -ov.first);
        }
    }
    g.n_works = (int64_t)hw.size();
    if (dev_upload(P, hw, &g.d_works)) return 1;
    if (dev_upload(P, hs, &g.d_segs)) return 1;
    P->n_launches += (int64_t)g.launches.size();
    return 0;
}

static int build_axpy_group(qcm_plan_s* P, AxpyGroup& g, qcm_wave_desc const& wd)
{
    std::vector<DWSrc> hs((size_t)wd.n_w_srcs);
    parallel_ranges(wd.n_w_srcs, 1 << 17, [&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) hs[i] = DWSrc{wd.w_srcs[i].src.off, wd.w_srcs[i].src.buf, wd.w_srcs[i].lds}; });
    std::vector<DWDst> hd((size_t)wd.n_w_dsts);
    parallel_ranges(wd.n_w_dsts, 1 << 17, [&](int, int64_t b, int64_t e) { for (int64_t i = b; i < e; ++i) hd[i] = DWDst{wd.w_dsts[i].dst.off, wd.w_dsts[i].dst.buf, wd.w_dsts[i].ldd}; });
    std::vector<DWGroup> hg((size_t)wd.n_w_groups);
    std::vector<DWWork> cls[5];
    for (int64_t i = 0; i < wd.n_w_groups; ++i) {
        qcm_w_group const& q = wd.w_groups[i];
        int c;
        if (q.cls == 1) {
            if (q.n_src > 4 || q.n_dst > 4 || q.ng != 4) return fail("qcm_plan_create: stream W group outside (n_src <= 4, n_dst <= 4, ng = 4)");
            c = 0;
        } else {
            c = q.ng == 8 ? 1 : q.ng == 16 ? 2 : q.ng == 32 ? 3 : q.ng == 64 ? 4 : -1;
            if (c < 0) return fail("qcm_plan_create: W group with ng outside {8,16,32,64}");
            if (q.n_dst > q.ng) return fail("qcm_plan_create: W group with more destinations than ng");
        }
        if (q.n_src < 1 || q.n_dst < 1) return fail("qcm_plan_create: empty W group");
        if ((int64_t)q.rows * q.cols >= ((int64_t)1 << 31)) return fail("qcm_plan_create: W panel with 2^31 or more elements");
        hg[i] = DWGroup{q.rows, q.cols, q.n_src, q.n_dst, q.ng, q.src_begin, q.dst_begin, q.cls, q.coef_begin};
        const int n = q.rows * q.cols, tile = c == 0 ? WS_TILE : wgemm_ws_tile();
        for (int e0 = 0; e0 < n; e0 += tile) cls[c].push_back(DWWork{(int)i, e0});
    }
    std::vector<DWWork> hw;
    for (int c = 0; c < 5; ++c) {
        g.begin[c] = (int64_t)hw.size(); g.count[c] = (int64_t)cls[c].size();
        hw.insert(hw.end(), cls[c].begin(), cls[c].end());
        if (g.count[c]) P->n_launches += 1;
    }
    if (dev_upload(P, hw, &g.d_works) || dev_upload(P, hg, &g.d_groups) || dev_upload(P, hs, &g.d_srcs) || dev_upload(P, hd, &g.d_dsts) ||
        dev_upload_raw(P, wd.w_coefs, (size_t)wd.n_w_coefs * sizeof(double), (void**)&g.d_coefs)) return 1;
    return 0;
}

static int wgemm_set_attributes()
{
    const char* e = wgemm_ws_init(G.sm_count);
    if (e) return fail(std::string("wgemm_ws_init: ") + e);
    return 0;
}
static int run_w_group(AxpyGroup const& g, BufTable const& bufs)
{
    // the classes write disjoint destination panels: run them side by side
    int n_active = 0;
    for (int c = 0; c < 5; ++c) n_active += g.count[c] > 0;
    if (n_active == 0) return 0;
    const bool fork = n_active > 1;
    if (fork) {
        CU(cudaEventRecord(G.fork_ev, G.stream));
        for (int i = 0; i < Global::kAux; ++i) CU(cudaStreamWaitEvent(G.aux[i], G.fork_ev, 0));
    }
    bool used[Global::kAux] = {false, false, false};
    int li = 0;
    for (int c = 4; c >= 0; --c) {       // the compute-heavy classes first
        if (!g.count[c]) continue;
        cudaStream_t st = G.stream;
        if (fork && li > 0) { int a = (li - 1) % Global::kAux; st = G.aux[a]; used[a] = true; }
        switch (c) {
            case 0: k_wstream<<<(unsigned)g.count[0], WS_THREADS, 0, st>>>(g.d_works + g.begin[0], g.d_groups, g.d_srcs, g.d_dsts, g.d_coefs, bufs); break;
            default: wgemm_ws_launch(c - 1, g.count[c], g.d_works + g.begin[c], g.d_groups, g.d_srcs, g.d_dsts, g.d_coefs, bufs, st); break;
        }
        G.launches++; ++li;
    }
    if (fork)
        for (int i = 0; i < Global::kAux; ++i)
            if (used[i]) { CU(cudaEventRecord(G.join_ev[i], G.aux[i])); CU(cudaStreamWaitEvent(G.stream, G.join_ev[i], 0)); }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int qcm_plan_create(const qcm_plan_desc* d, qcm_plan_t* out)
{
    CHECK_INIT();
    if (!d || !out) return fail("qcm_plan_create: null argument");
    static std::mutex plan_mutex;      // the pinned staging buffer is shared
    std::lock_guard<std::mutex> lock(plan_mutex);
    qcm_plan_s* P = new qcm_plan_s();
    g_pin.used = 0;
    P->kind = d->kind; P->flops = d->flops; P->bytes = d->bytes;
    P->world = d->world > 1 ? d->world : 1; P->rank = d->world > 1 ? d->rank : 0;
    if (P->rank < 0 || P->rank >= P->world) { delete P; return fail("qcm_plan_create: rank outside [0, world)"); }
    for (int i = 0; i < QCM_BUF_COUNT; ++i) P->elems[i] = d->elems[i];
    auto bail = [&]() { for (void* p : P->allocs) cudaFreeAsync(p, G.stream); delete P; return 1; };
    {
        std::vector<DCopy> hc((size_t)d->n_pre_copies);
        for (int64_t i = 0; i < d->n_pre_copies; ++i) {
            qcm_copy_task const& c = d->pre_copies[i];
            hc[i] = DCopy{c.src.off, c.dst.off, c.src.buf, c.dst.buf, c.rows, c.cols, c.lds, c.ldd};
        }
        P->n_copies = d->n_pre_copies;
        if (dev_upload(P, hc, &P->d_copies)) return bail();
        if (P->n_copies) P->n_launches++;
    }
    if (build_gemm_group(P, P->p, d->p_outs, d->n_p_outs, d->p_segs, d->n_p_segs, 0)) return bail();
    P->waves.resize(d->n_waves);
    for (int w = 0; w < d->n_waves; ++w) {
        qcm_wave_desc const& wd = d->waves[w];
        WaveDev& W = P->waves[w];
        W.y_elems = wd.y_elems; W.t_elems = wd.t_elems; W.x_chunk = wd.x_chunk_elems; W.x_zero = wd.x_zero != 0;
        if (W.x_chunk < 0 || (W.x_chunk > 0 && (w != 0 || P->world <= 1))) return (fail("qcm_plan_create: an exchange wave must be waves[0] of a sharded plan"), bail());
        if (W.x_chunk * P->world > P->elems[QCM_BUF_Y]) return (fail("qcm_plan_create: exchange region larger than the Y workspace"), bail());
        if (build_gemm_group(P, W.t, wd.t_outs, wd.n_t_outs, wd.t_segs, wd.n_t_segs, 0)) return bail();
        if (build_axpy_group(P, W.w, wd)) return bail();
        if (build_gemm_group(P, W.c, wd.c_outs, wd.n_c_outs, wd.c_segs, wd.n_c_segs, 1)) return bail();
    }
    if (flush_uploads(P)) return bail();
    *out = P;
    return 0;
}

extern "C" int qcm_plan_destroy(qcm_plan_t P)
{
    if (!P) return 0;
    qcm_internal_forget_plan(P);
    // the task arrays are released in stream order: kernels of this plan that are still queued finish first
    for (size_t i = 0; i < P->allocs.size(); ++i) {
        void* p = P->allocs[i];
        const size_t cap = i < P->alloc_caps.size() ? P->alloc_caps[i] : 0;
        if (G.ready && cap > 0) {
            g_task_bufs.push_back(std::make_pair((char*)p, cap));
            if (g_task_bufs.size() > kTaskBufsKept) {       // drop the smallest
                size_t m = 0;
                for (size_t j = 1; j < g_task_bufs.size(); ++j) if (g_task_bufs[j].second < g_task_bufs[m].second) m = j;
                cudaFreeAsync(g_task_bufs[m].first, G.stream);
                g_task_bufs.erase(g_task_bufs.begin() + (long)m);
            }
        } else if (G.ready) cudaFreeAsync(p, G.stream); else cudaFree(p);
    }
    delete P;
    return 0;
}

extern "C" int qcm_plan_stats(qcm_plan_t P, double* flops, int64_t* bytes, int64_t* n_launches, int64_t* workspace_bytes)
{
    if (!P) return fail("null plan");
    if (flops) *flops = P->flops;
    if (bytes) *bytes = P->bytes;
    if (n_launches) *n_launches = P->n_launches;
    if (workspace_bytes) {
        int64_t s = 0;
        for (int slot : {QCM_BUF_KET_RP, QCM_BUF_T, QCM_BUF_TP, QCM_BUF_Y, QCM_BUF_BRA_RP}) s += P->elems[slot];
        *workspace_bytes = s * 8;
    }
    return 0;
}

// ---- execution ------------------------------------------------------------------------------------------
static int run_gemm_group(GemmGroup const& g, BufTable const& bufs)
{
    if (g.launches.empty()) return 0;
    // the variants of one group write disjoint output tiles (or combine with atomics): they may overlap, which
    // hides the tails of the small-tile launches behind the large ones
    const bool fork = g.launches.size() > 1;
    if (fork) {
        CU(cudaEventRecord(G.fork_ev, G.stream));
        for (int i = 0; i < Global::kAux; ++i) CU(cudaStreamWaitEvent(G.aux[i], G.fork_ev, 0));
    }
    int li = 0;
    bool used[Global::kAux] = {false, false, false};
    for (auto const& l : g.launches) {
        cudaStream_t st = G.stream;
        if (fork && li > 0) { int a = (li - 1) % Global::kAux; st = G.aux[a]; used[a] = true; }
        gemm_ws_launch(l.variant, l.n_works, g.d_works + l.work_begin, g.d_segs, bufs, st);
        G.launches++; ++li;
    }
    if (fork)
        for (int i = 0; i < Global::kAux; ++i)
            if (used[i]) { CU(cudaEventRecord(G.join_ev[i], G.aux[i])); CU(cudaStreamWaitEvent(G.stream, G.join_ev[i], 0)); }
    CU(cudaGetLastError());
    return 0;
}

// bufs must have every input/output slot bound; workspaces are bound here
static int execute(qcm_plan_s* P, BufTable bufs)
{
    for (int slot : {QCM_BUF_KET_RP, QCM_BUF_T, QCM_BUF_TP, QCM_BUF_Y, QCM_BUF_BRA_RP}) {
        if (bufs.p[slot]) continue;
        if (ensure_ws(slot, P->elems[slot])) return 1;
        bufs.p[slot] = G.ws[slot];
    }
    bool tm = G.timing;
    float acc[4] = {0, 0, 0, 0}, x_wait_ms = 0;
    auto mark = [&](int i) { if (tm) cudaEventRecord(G.ev[i], G.stream); };
    auto lap = [&](int phase, int i0, int i1) { if (tm) { cudaEventSynchronize(G.ev[i1]); float ms = 0; cudaEventElapsedTime(&ms, G.ev[i0], G.ev[i1]); acc[phase] += ms; } };
    mark(0);
    if (P->n_copies) {
        // right-paired tensors have unused gaps only if a sector is absent; zero them once per call
        if (P->kind != 2 && P->elems[QCM_BUF_KET_RP]) CU(cudaMemsetAsync(bufs.p[QCM_BUF_KET_RP], 0, (size_t)P->elems[QCM_BUF_KET_RP] * 8, G.stream));
        if (P->kind == 2 && P->elems[QCM_BUF_BRA_RP]) CU(cudaMemsetAsync(bufs.p[QCM_BUF_BRA_RP], 0, (size_t)P->elems[QCM_BUF_BRA_RP] * 8, G.stream));
        k_copy_panels<<<(unsigned)P->n_copies, 128, 0, G.stream>>>(P->d_copies, bufs);
        G.launches++;
    }
    mark(1); lap(0, 0, 1);
    if (run_gemm_group(P->p, bufs)) return 1;
    mark(2); lap(1, 1, 2);
    // QCM_SYNC_DEBUG: synchronise and report after every phase (localises a stuck kernel or collective)
    static const bool dbg = getenv("QCM_SYNC_DEBUG") != nullptr;
    auto checkpoint = [&](const char* what) -> int {
        if (!dbg) return 0;
        cudaError_t e1 = cudaStreamSynchronize(G.stream);
        fprintf(stderr, "[qcm rank %d] %s: %s\n", G.rank, what, cudaGetErrorString(e1)); fflush(stderr);
        return e1 == cudaSuccess ? 0 : fail(std::string(what) + ": " + cudaGetErrorString(e1));
    };
    if (checkpoint("reshape + resident step-1 products")) return 1;
    const WaveDev* xwave = nullptr;
    for (auto const& W : P->waves) {
        mark(3);
        if (W.x_chunk > 0) {
            // exchange wave: partial sums -> exchange region -> reduce-scatter on the communication stream, overlapped with
            // the waves that follow; its closing products run once the complete sums of this rank's chunk have arrived
            if (!G.comm || G.world != P->world) return fail("exchange wave without a matching communicator");
            if (W.x_zero) CU(cudaMemsetAsync(bufs.p[QCM_BUF_Y], 0, (size_t)W.x_chunk * P->world * 8, G.stream));
            if (run_w_group(W.w, bufs)) return 1;
            if (checkpoint("exchange wave: W pass")) return 1;
            CU(cudaEventRecord(G.x_ready, G.stream));
            CU(cudaStreamWaitEvent(G.comm_stream, G.x_ready, 0));
            int r = G.f_rs(bufs.p[QCM_BUF_Y], bufs.p[QCM_BUF_Y] + (size_t)P->rank * W.x_chunk, (size_t)W.x_chunk, 8 /*ncclFloat64*/, 0 /*ncclSum*/, G.comm, G.comm_stream);
            if (r != 0) return fail(std::string("ncclReduceScatter: ") + (G.f_err ? G.f_err(r) : "error"));
            CU(cudaEventRecord(G.x_done, G.comm_stream));
            if (dbg) { cudaError_t e2 = cudaStreamSynchronize(G.comm_stream); fprintf(stderr, "[qcm rank %d] reduce-scatter of %lld x %d elements: %s\n", G.rank, (long long)W.x_chunk, P->world, cudaGetErrorString(e2)); fflush(stderr); }
            xwave = &W;
            mark(5); lap(2, 3, 5);
            continue;
        }
        if (run_gemm_group(W.t, bufs)) return 1;
        mark(4); lap(1, 3, 4);
        if (run_w_group(W.w, bufs)) return 1;
        mark(5); lap(2, 4, 5);
        if (run_gemm_group(W.c, bufs)) return 1;
        mark(6); lap(3, 5, 6);
        if (checkpoint("local wave")) return 1;
    }
    if (xwave) {
        mark(3);
        CU(cudaStreamWaitEvent(G.stream, G.x_done, 0));
        mark(4);
        if (run_gemm_group(xwave->c, bufs)) return 1;
        if (checkpoint("closing products of the exchange chunk")) return 1;
        mark(5); lap(3, 4, 5);
        if (tm) { cudaEventSynchronize(G.ev[4]); float ms = 0; cudaEventElapsedTime(&ms, G.ev[3], G.ev[4]); x_wait_ms = ms; }
    }
    CU(cudaGetLastError());
    if (tm) { for (int i = 0; i < 4; ++i) G.last_ms[i] = acc[i]; G.last_ms[4] = x_wait_ms; G.last_ms[5] = acc[0] + acc[1] + acc[2] + acc[3] + x_wait_ms; }
    return 0;
}

// Time-sliced shards.  A site problem whose resident step-1 products do not fit one device (cfg4: 145 GB) is planned as V
// shards -- exactly the plans V ranks would run (edges of the MPO bond graph sharded by their step-1 index) -- and the shards are
// executed one after another on this device: only one shard's step-1 products are resident at a time.  The exchange of
// partial W sums becomes a local accumulation (the region of every shard is added into an accumulator; the closing products
// of all chunks then read the complete sums), the allreduce of sigma becomes accumulation into the same output.
static int execute_sliced(qcm_plan_s* const* Ps, int V, BufTable bufs)
{
    for (int slot : {QCM_BUF_KET_RP, QCM_BUF_T, QCM_BUF_TP, QCM_BUF_Y, QCM_BUF_BRA_RP}) {
        int64_t need = 0;
        for (int v = 0; v < V; ++v) need = std::max(need, Ps[v]->elems[slot]);
        if (ensure_ws(slot, need)) return 1;
        bufs.p[slot] = G.ws[slot];
    }
    int64_t x_elems = 0;
    for (int v = 0; v < V; ++v) {
        qcm_plan_s* P = Ps[v];
        const int64_t xe = (!P->waves.empty() && P->waves[0].x_chunk > 0) ? P->waves[0].x_chunk * P->world : 0;
        if (v == 0) x_elems = xe; else if (xe != x_elems) return fail("qcm_site_hamil2_sliced: the shards disagree on the exchange region");
    }
    if (x_elems > G.xacc_elems) {
        if (G.xacc) { CU(cudaStreamSynchronize(G.stream)); CU(cudaFree(G.xacc)); G.xacc = nullptr; G.xacc_elems = 0; }
        CU(cudaMalloc((void**)&G.xacc, (size_t)x_elems * 8));
        G.xacc_elems = x_elems;
    }
    const bool tm = G.timing;
    float acc[4] = {0, 0, 0, 0};
    auto mark = [&](int i) { if (tm) cudaEventRecord(G.ev[i], G.stream); };
    auto lap = [&](int phase, int i0, int i1) { if (tm) { cudaEventSynchronize(G.ev[i1]); float ms = 0; cudaEventElapsedTime(&ms, G.ev[i0], G.ev[i1]); acc[phase] += ms; } };
    for (int v = 0; v < V; ++v) {
        qcm_plan_s* P = Ps[v];
        mark(0);
        if (v == 0 && P->n_copies) {
            if (P->elems[QCM_BUF_KET_RP]) CU(cudaMemsetAsync(bufs.p[QCM_BUF_KET_RP], 0, (size_t)P->elems[QCM_BUF_KET_RP] * 8, G.stream));
            k_copy_panels<<<(unsigned)P->n_copies, 128, 0, G.stream>>>(P->d_copies, bufs);
            G.launches++;
        }
        mark(1); lap(0, 0, 1);
        if (run_gemm_group(P->p, bufs)) return 1;
        mark(2); lap(1, 1, 2);
        for (auto const& W : P->waves) {
            mark(3);
            if (W.x_chunk > 0) {
                if (W.x_zero) CU(cudaMemsetAsync(bufs.p[QCM_BUF_Y], 0, (size_t)x_elems * 8, G.stream));
                if (run_w_group(W.w, bufs)) return 1;
                if (v == 0) CU(cudaMemcpyAsync(G.xacc, bufs.p[QCM_BUF_Y], (size_t)x_elems * 8, cudaMemcpyDeviceToDevice, G.stream));
                else { k_vec_axpy<<<G.sm_count * 8, 256, 0, G.stream>>>(1.0, bufs.p[QCM_BUF_Y], G.xacc, x_elems); G.launches++; }
                mark(4); lap(2, 3, 4);
                continue;
            }
            if (run_gemm_group(W.t, bufs)) return 1;
            mark(4); lap(1, 3, 4);
            if (run_w_group(W.w, bufs)) return 1;
            mark(5); lap(2, 4, 5);
            if (run_gemm_group(W.c, bufs)) return 1;
            mark(6); lap(3, 5, 6);
        }
    }
    if (x_elems > 0) {
        mark(3);
        CU(cudaMemcpyAsync(bufs.p[QCM_BUF_Y], G.xacc, (size_t)x_elems * 8, cudaMemcpyDeviceToDevice, G.stream));
        for (int v = 0; v < V; ++v) if (run_gemm_group(Ps[v]->waves[0].c, bufs)) return 1;
        mark(4); lap(3, 3, 4);
    }
    if (tm) { for (int i = 0; i < 4; ++i) G.last_ms[i] = acc[i]; G.last_ms[4] = 0; G.last_ms[5] = acc[0] + acc[1] + acc[2] + acc[3]; }
    CU(cudaGetLastError());
    return 0;
}

static int check_arr(qcm_array_t a, int64_t need, const char* what)
{
    if (!a) return fail(std::string(what) + ": null array");
    if (a->n < need) return fail(std::string(what) + ": array holds " + std::to_string(a->n) + " elements, plan needs " + std::to_string(need));
    if (!a->resident) return fail(std::string(what) + ": the array has been evicted to host memory (qcm_array_prefetch brings it back)");
    if (a->pending) { CU(cudaStreamWaitEvent(G.stream, a->ready, 0)); a->pending = false; }     // a prefetch is in flight: later work waits for it
    return 0;
}

static int allreduce_ptr(double* p, int64_t n);
// a plan sharded over `world` ranks produces a partial result: it may only run under a communicator of that shape
static int check_sharding(qcm_plan_s const* P)
{
    if (P->world <= 1) return 0;
    if (P->world != G.world || P->rank != G.rank)
        return fail("plan was built for rank " + std::to_string(P->rank) + " of " + std::to_string(P->world) + " but the communicator is rank " +
                    std::to_string(G.rank) + " of " + std::to_string(G.world) + " (qcm_comm_init)");
    return 0;
}

extern "C" int qcm_site_hamil2_dev(qcm_plan_t P, qcm_array_t left, qcm_array_t right, qcm_array_t psi, qcm_array_t sigma)
{
    CHECK_INIT();
    if (!P || P->kind != 0) return fail("qcm_site_hamil2: plan is not a sigma plan");
    if (check_sharding(P)) return 1;
    if (check_arr(left, P->elems[QCM_BUF_LEFT], "left boundary") || check_arr(right, P->elems[QCM_BUF_RIGHT], "right boundary") ||
        check_arr(psi, P->elems[QCM_BUF_KET_LP], "psi") || check_arr(sigma, P->elems[QCM_BUF_OUT], "sigma")) return 1;
    BufTable b; memset(&b, 0, sizeof(b));
    b.p[QCM_BUF_LEFT] = left->p; b.p[QCM_BUF_RIGHT] = right->p; b.p[QCM_BUF_KET_LP] = psi->p; b.p[QCM_BUF_OUT] = sigma->p;
    if (P->elems[QCM_BUF_OUT]) CU(cudaMemsetAsync(sigma->p, 0, (size_t)P->elems[QCM_BUF_OUT] * 8, G.stream));
    if (execute(P, b)) return 1;
    if (P->world > 1) return allreduce_ptr(sigma->p, P->elems[QCM_BUF_OUT]);
    return 0;
}

extern "C" int qcm_site_hamil2_sliced_dev(const qcm_plan_t* plans, int n, qcm_array_t left, qcm_array_t right, qcm_array_t psi, qcm_array_t sigma)
{
    CHECK_INIT();
    if (!plans || n < 1) return fail("qcm_site_hamil2_sliced: no plans");
    for (int v = 0; v < n; ++v) {
        qcm_plan_s* P = plans[v];
        if (!P || P->kind != 0) return fail("qcm_site_hamil2_sliced: plan is not a sigma plan");
        if ((n > 1 && (P->world != n || P->rank != v)) || (n == 1 && P->world != 1)) return fail("qcm_site_hamil2_sliced: plans[v] must be shard v of n");
        if (check_arr(left, P->elems[QCM_BUF_LEFT], "left boundary") || check_arr(right, P->elems[QCM_BUF_RIGHT], "right boundary") ||
            check_arr(psi, P->elems[QCM_BUF_KET_LP], "psi") || check_arr(sigma, P->elems[QCM_BUF_OUT], "sigma")) return 1;
    }
    BufTable b; memset(&b, 0, sizeof(b));
    b.p[QCM_BUF_LEFT] = left->p; b.p[QCM_BUF_RIGHT] = right->p; b.p[QCM_BUF_KET_LP] = psi->p; b.p[QCM_BUF_OUT] = sigma->p;
    if (plans[0]->elems[QCM_BUF_OUT]) CU(cudaMemsetAsync(sigma->p, 0, (size_t)plans[0]->elems[QCM_BUF_OUT] * 8, G.stream));
    return execute_sliced(plans, n, b);
}
extern "C" int qcm_site_hamil2_sliced(const qcm_plan_t* plans, int n, qcm_array_t left, qcm_array_t right, const double* psi, double* sigma)
{
    CHECK_INIT();
    if (!plans || n < 1 || !plans[0]) return fail("qcm_site_hamil2_sliced: no plans");
    qcm_plan_s* P = plans[0];
    if (ensure_ws(QCM_BUF_KET_LP, P->elems[QCM_BUF_KET_LP]) || ensure_ws(QCM_BUF_OUT, P->elems[QCM_BUF_OUT])) return 1;
    qcm_array_s a_psi, a_sig;
    a_psi.p = G.ws[QCM_BUF_KET_LP]; a_psi.n = G.ws_elems[QCM_BUF_KET_LP]; a_sig.p = G.ws[QCM_BUF_OUT]; a_sig.n = G.ws_elems[QCM_BUF_OUT];
    if (P->elems[QCM_BUF_KET_LP]) CU(cudaMemcpyAsync(a_psi.p, psi, (size_t)P->elems[QCM_BUF_KET_LP] * 8, cudaMemcpyHostToDevice, G.stream));
    if (qcm_site_hamil2_sliced_dev(plans, n, left, right, &a_psi, &a_sig)) return 1;
    if (P->elems[QCM_BUF_OUT]) CU(cudaMemcpyAsync(sigma, a_sig.p, (size_t)P->elems[QCM_BUF_OUT] * 8, cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}

extern "C" int qcm_site_hamil2(qcm_plan_t P, qcm_array_t left, qcm_array_t right, const double* psi, double* sigma)
{
    CHECK_INIT();
    if (!P || P->kind != 0) return fail("qcm_site_hamil2: plan is not a sigma plan");
    if (ensure_ws(QCM_BUF_KET_LP, P->elems[QCM_BUF_KET_LP]) || ensure_ws(QCM_BUF_OUT, P->elems[QCM_BUF_OUT])) return 1;
    qcm_array_s a_psi, a_sig;
    a_psi.p = G.ws[QCM_BUF_KET_LP]; a_psi.n = G.ws_elems[QCM_BUF_KET_LP]; a_sig.p = G.ws[QCM_BUF_OUT]; a_sig.n = G.ws_elems[QCM_BUF_OUT];
    if (P->elems[QCM_BUF_KET_LP]) CU(cudaMemcpyAsync(a_psi.p, psi, (size_t)P->elems[QCM_BUF_KET_LP] * 8, cudaMemcpyHostToDevice, G.stream));
    if (qcm_site_hamil2_dev(P, left, right, &a_psi, &a_sig)) return 1;
    if (P->elems[QCM_BUF_OUT]) CU(cudaMemcpyAsync(sigma, a_sig.p, (size_t)P->elems[QCM_BUF_OUT] * 8, cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}

extern "C" int qcm_boundary_step(qcm_plan_t P, qcm_array_t in, const double* bra, const double* ket, qcm_array_t out)
{
    CHECK_INIT();
    if (!P || (P->kind != 1 && P->kind != 2)) return fail("qcm_boundary_step: plan is not a boundary plan");
    int in_slot = P->kind == 1 ? QCM_BUF_LEFT : QCM_BUF_RIGHT;
    if (check_sharding(P)) return 1;
    if (check_arr(in, P->elems[in_slot], "input boundary") || check_arr(out, P->elems[QCM_BUF_OUT], "output boundary")) return 1;
    if (ensure_ws(QCM_BUF_KET_LP, P->elems[QCM_BUF_KET_LP]) || ensure_ws(QCM_BUF_BRA_LP, P->elems[QCM_BUF_BRA_LP])) return 1;
    if (P->elems[QCM_BUF_KET_LP]) CU(cudaMemcpyAsync(G.ws[QCM_BUF_KET_LP], ket, (size_t)P->elems[QCM_BUF_KET_LP] * 8, cudaMemcpyHostToDevice, G.stream));
    if (P->elems[QCM_BUF_BRA_LP]) CU(cudaMemcpyAsync(G.ws[QCM_BUF_BRA_LP], bra, (size_t)P->elems[QCM_BUF_BRA_LP] * 8, cudaMemcpyHostToDevice, G.stream));
    BufTable b; memset(&b, 0, sizeof(b));
    b.p[in_slot] = in->p; b.p[QCM_BUF_KET_LP] = G.ws[QCM_BUF_KET_LP]; b.p[QCM_BUF_BRA_LP] = G.ws[QCM_BUF_BRA_LP]; b.p[QCM_BUF_OUT] = out->p;
    if (out->n) CU(cudaMemsetAsync(out->p, 0, (size_t)out->n * 8, G.stream));
    if (execute(P, b)) return 1;
    // every rank computed its share of the output bond indices into a zeroed array: the sum is the full boundary
    if (P->world > 1 && allreduce_ptr(out->p, P->elems[QCM_BUF_OUT])) return 1;
    // The step is QUEUED, not awaited: `out` stays on the device and every later use of it (the next sigma, the next boundary step,
    // qcm_array_download) is ordered behind it on the library's stream, so the host goes on to plan the next site while the device
    // works.  (bra / ket were pageable host memory: cudaMemcpyAsync has staged them before returning.)
    if (G.timing) CU(cudaStreamSynchronize(G.stream));
    return 0;
}

extern "C" int qcm_hdiag(qcm_plan_t P, qcm_array_t left, qcm_array_t right, double* diag)
{
    CHECK_INIT();
    if (!P || P->kind != 3) return fail("qcm_hdiag: plan is not a diagonal_hamiltonian plan");
    if (check_arr(left, P->elems[QCM_BUF_LEFT], "left boundary") || check_arr(right, P->elems[QCM_BUF_RIGHT], "right boundary")) return 1;
    if (ensure_ws(QCM_BUF_OUT, P->elems[QCM_BUF_OUT]) || ensure_ws(QCM_BUF_Y, P->elems[QCM_BUF_Y])) return 1;
    BufTable b; memset(&b, 0, sizeof(b));
    b.p[QCM_BUF_LEFT] = left->p; b.p[QCM_BUF_RIGHT] = right->p; b.p[QCM_BUF_OUT] = G.ws[QCM_BUF_OUT]; b.p[QCM_BUF_Y] = G.ws[QCM_BUF_Y];
    // rows of V without contributions and the accumulating output start from zero
    if (P->elems[QCM_BUF_OUT]) CU(cudaMemsetAsync(b.p[QCM_BUF_OUT], 0, (size_t)P->elems[QCM_BUF_OUT] * 8, G.stream));
    if (P->elems[QCM_BUF_Y]) CU(cudaMemsetAsync(b.p[QCM_BUF_Y], 0, (size_t)P->elems[QCM_BUF_Y] * 8, G.stream));
    if (execute(P, b)) return 1;
    if (P->elems[QCM_BUF_OUT]) CU(cudaMemcpyAsync(diag, b.p[QCM_BUF_OUT], (size_t)P->elems[QCM_BUF_OUT] * 8, cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}

extern "C" int qcm_set_timing(int enabled) { G.timing = enabled != 0; return 0; }
extern "C" int qcm_last_timing(double ms[6]) { for (int i = 0; i < 6; ++i) ms[i] = G.last_ms[i]; return 0; }

// ---- BLAS-1 ---------------------------------------------------------------------------------------------
static int dot_blocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(DOT_MAX_BLOCKS, (n + DOT_THREADS * 8 - 1) / (DOT_THREADS * 8))); }
extern "C" int qcm_vec_dots(const qcm_array_t* xs, const qcm_array_t* ys, int k, int64_t n, double* results)
{
    CHECK_INIT();
    if (k < 0 || k > QCM_MAX_DOTS) return fail("qcm_vec_dots: between 0 and QCM_MAX_DOTS pairs per call");
    if (k == 0) return 0;
    for (int q = 0; q < k; ++q) if (check_arr(xs[q], n, "x") || check_arr(ys[q], n, "y")) return 1;
    if (!G.dot_partial) CU(cudaMalloc((void**)&G.dot_partial, (size_t)(QCM_MAX_DOTS * (DOT_MAX_BLOCKS + 1)) * sizeof(double)));
    const int nb = dot_blocks(n);
    double* out = G.dot_partial + (size_t)QCM_MAX_DOTS * DOT_MAX_BLOCKS;
    if (n == 0) { for (int q = 0; q < k; ++q) results[q] = 0.; return 0; }
    for (int q = 0; q < k; ++q) {
        k_vec_dot_partial<<<nb, DOT_THREADS, 0, G.stream>>>(xs[q]->p, ys[q]->p, n, G.dot_partial + (size_t)q * DOT_MAX_BLOCKS);
        G.launches++;
    }
    k_vec_dot_final<<<k, DOT_THREADS, 0, G.stream>>>(G.dot_partial, nb, DOT_MAX_BLOCKS, out);
    G.launches++;
    CU(cudaMemcpyAsync(results, out, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}
extern "C" int qcm_vec_dot(qcm_array_t x, qcm_array_t y, int64_t n, double* result) { return qcm_vec_dots(&x, &y, 1, n, result); }
// out = sum_j coefs[j] * xs[j]   (out may not alias any xs[j])
__global__ void k_vec_lincomb(const double* const* __restrict__ xs, const double* __restrict__ coefs, int k, double* __restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = 0.;
        for (int j = 0; j < k; ++j) s = fma(coefs[j], xs[j][i], s);
        out[i] = s;
    }
}
extern "C" int qcm_vec_lincomb(const qcm_array_t* xs, const double* coefs, int k, qcm_array_t out, int64_t n)
{
    CHECK_INIT();
    if (k < 1 || k > QCM_MAX_DOTS) return fail("qcm_vec_lincomb: between 1 and QCM_MAX_DOTS terms per call");
    if (check_arr(out, n, "out")) return 1;
    for (int q = 0; q < k; ++q) { if (check_arr(xs[q], n, "x")) return 1; if (xs[q] == out) return fail("qcm_vec_lincomb: out aliases an input"); }
    if (n == 0) return 0;
    // arguments travel through a small device area; the previous call's kernel may still read it: one slot per call, round robin
    if (!G.lc_args) CU(cudaMalloc((void**)&G.lc_args, (size_t)Global::kLcSlots * QCM_MAX_DOTS * 16));
    char* slot = G.lc_args + (size_t)(G.lc_next++ % Global::kLcSlots) * QCM_MAX_DOTS * 16;
    const double* hp[QCM_MAX_DOTS]; double hc[QCM_MAX_DOTS];
    for (int q = 0; q < k; ++q) { hp[q] = xs[q]->p; hc[q] = coefs[q]; }
    if (G.lc_next % Global::kLcSlots == 0) CU(cudaStreamSynchronize(G.stream));      // slots are reused only after the stream drained
    CU(cudaMemcpyAsync(slot, hp, (size_t)k * 8, cudaMemcpyHostToDevice, G.stream));
    CU(cudaMemcpyAsync(slot + QCM_MAX_DOTS * 8, hc, (size_t)k * 8, cudaMemcpyHostToDevice, G.stream));
    k_vec_lincomb<<<std::max(1, std::min(G.sm_count * 8, (int)((n + 255) / 256))), 256, 0, G.stream>>>((const double* const*)slot, (const double*)(slot + QCM_MAX_DOTS * 8), k, out->p, n);
    G.launches++;
    CU(cudaGetLastError());
    return 0;
}
extern "C" int qcm_vec_axpy(double a, qcm_array_t x, qcm_array_t y, int64_t n)
{
    CHECK_INIT();
    if (check_arr(x, n, "x") || check_arr(y, n, "y")) return 1;
    if (n) { k_vec_axpy<<<std::max(1, std::min(G.sm_count * 8, (int)((n + 255) / 256))), 256, 0, G.stream>>>(a, x->p, y->p, n); G.launches++; }
    CU(cudaGetLastError());
    return 0;
}
extern "C" int qcm_vec_scal(double a, qcm_array_t x, int64_t n)
{
    CHECK_INIT();
    if (check_arr(x, n, "x")) return 1;
    if (n) { k_vec_scal<<<std::max(1, std::min(G.sm_count * 8, (int)((n + 255) / 256))), 256, 0, G.stream>>>(a, x->p, n); G.launches++; }
    CU(cudaGetLastError());
    return 0;
}
extern "C" int qcm_vec_copy(qcm_array_t src, qcm_array_t dst, int64_t n)
{
    CHECK_INIT();
    if (check_arr(src, n, "src") || check_arr(dst, n, "dst")) return 1;
    if (n) CU(cudaMemcpyAsync(dst->p, src->p, (size_t)n * 8, cudaMemcpyDeviceToDevice, G.stream));
    return 0;
}

// ---- NCCL ----------------------------------------------------------------------------------------------
static int nccl_resolve()
{
    if (G.f_ar) return 0;
    void* h = nullptr;
    // prefer a libnccl that is already loaded into the process (e.g. the one bundled with PyTorch)
    if (dlsym(RTLD_DEFAULT, "ncclAllReduce")) h = RTLD_DEFAULT;
    else {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(std::string("cannot load libnccl.so.2: ") + dlerror());
        G.nccl_lib = h;
    }
    G.f_uid = (nccl_get_uid_fn)dlsym(h, "ncclGetUniqueId");
    G.f_init = (nccl_init_rank_fn)dlsym(h, "ncclCommInitRank");
    G.f_ar = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    G.f_rs = (nccl_reducescatter_fn)dlsym(h, "ncclReduceScatter");
    G.f_destroy = (nccl_destroy_fn)dlsym(h, "ncclCommDestroy");
    G.f_err = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    if (!G.f_uid || !G.f_init || !G.f_ar || !G.f_rs || !G.f_destroy) { G.f_ar = nullptr; return fail("libnccl lacks required symbols"); }
    return 0;
}
extern "C" int qcm_comm_unique_id(char id[128])
{
    if (nccl_resolve()) return 1;
    NcclUid u; memset(&u, 0, sizeof(u));
    int r = G.f_uid(&u);
    if (r != 0) return fail(std::string("ncclGetUniqueId: ") + (G.f_err ? G.f_err(r) : "error"));
    memcpy(id, u.internal, 128);
    return 0;
}
extern "C" int qcm_comm_init(int rank, int world, const char id[128])
{
    CHECK_INIT();
    if (world <= 1) { G.rank = 0; G.world = 1; return 0; }
    if (nccl_resolve()) return 1;
    NcclUid u; memcpy(u.internal, id, 128);
    int r = G.f_init(&G.comm, world, u, rank);
    if (r != 0) return fail(std::string("ncclCommInitRank: ") + (G.f_err ? G.f_err(r) : "error"));
    G.rank = rank; G.world = world;
    return 0;
}
extern "C" int qcm_comm_destroy(void)
{
    if (G.comm) { G.f_destroy(G.comm); G.comm = nullptr; }
    G.rank = 0; G.world = 1;
    return 0;
}
static int allreduce_ptr(double* p, int64_t n)
{
    if (G.world <= 1 || n == 0) return 0;
    if (!G.comm) return fail("allreduce without communicator");
    if (G.timing) cudaEventRecord(G.ev[6], G.stream);
    int r = G.f_ar(p, p, (size_t)n, 8 /*ncclFloat64*/, 0 /*ncclSum*/, G.comm, G.stream);
    if (r != 0) return fail(std::string("ncclAllReduce: ") + (G.f_err ? G.f_err(r) : "error"));
    if (G.timing) {
        cudaEventRecord(G.ev[7], G.stream); cudaEventSynchronize(G.ev[7]);
        float ms = 0; cudaEventElapsedTime(&ms, G.ev[6], G.ev[7]); G.last_ms[4] += ms; G.last_ms[5] += ms;
    }
    return 0;
}
extern "C" int qcm_comm_allreduce(qcm_array_t a, int64_t n)
{
    CHECK_INIT();
    if (check_arr(a, n, "allreduce")) return 1;
    return allreduce_ptr(a->p, n);
}

// ---- peak probes ----------------------------------------------------------------------------------------
template <class F> static int time_ms(F f, float* ms)
{
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    f();   // warm-up
    CU(cudaEventRecord(a, G.stream));
    f();
    CU(cudaEventRecord(b, G.stream));
    CU(cudaEventSynchronize(b));
    CU(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    CU(cudaGetLastError());
    return 0;
}
extern "C" int qcm_measure_fp64_fma_peak(double* tflops)
{
    CHECK_INIT();
    const int iters = 20000, blocks = G.sm_count * 8, threads = 256;
    float ms = 0;
    if (time_ms([&]() { k_peak_fma<<<blocks, threads, 0, G.stream>>>(G.scratch, iters); }, &ms)) return 1;
    *tflops = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    return 0;
}
extern "C" int qcm_measure_fp64_dmma_peak(double* tflops)
{
    CHECK_INIT();
    const int iters = 20000, blocks = G.sm_count * 8, threads = 256;
    float ms = 0;
    if (time_ms([&]() { k_peak_dmma<<<blocks, threads, 0, G.stream>>>(G.scratch, iters); }, &ms)) return 1;
    *tflops = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / (ms * 1e-3) / 1e12;
    return 0;
}
extern "C" int qcm_measure_hbm_copy(double* gbs)
{
    CHECK_INIT();
    const long long n = 1ll << 28;   // 2 GiB in + 2 GiB out, far beyond L2
    double *s = nullptr, *d = nullptr;
    CU(cudaMalloc((void**)&s, n * 8)); CU(cudaMalloc((void**)&d, n * 8));
    CU(cudaMemsetAsync(s, 0, n * 8, G.stream));
    float ms = 0, best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        if (time_ms([&]() { k_copy_stream<<<G.sm_count * 16, 512, 0, G.stream>>>((const double2*)s, (double2*)d, n / 2); }, &ms)) { cudaFree(s); cudaFree(d); return 1; }
        best = std::min(best, ms);
    }
    cudaFree(s); cudaFree(d);
    *gbs = 2.0 * n * 8 / (best * 1e-3) / 1e9;
    return 0;
}
