// qcm_b200.cu -- sm_100a execution layer behind include/qcm_b200.h.
//
// What runs here replaces the dense inner loops of QCMaquis's contraction engine
// (dmrg/mp_tensors/contractions/{abelian,non-abelian,common}): the block dgemm calls of
// block_matrix_algorithms.h:48-162 / non-abelian/gemm.hpp:48-204, the W-application axpy panels of
// alps_detail.hpp:189-224 / micro_kernels.hpp:19-198, the pairing reshapes of reshapes.h:177-335 and the
// critical-section reduction over the MPO bond index (abelian/site_hamil.hpp:84-86).  The host flattens
// those loops into task arrays once per (site, direction); the kernels below execute them:
//
//   k_copy_panels   pairing reshapes as panel copies
//   k_gemm_dmma     grouped, variable-size FP64 GEMM.  One CTA = one output tile; it walks a list of
//                   K-segments (the sum over MPO bond terms that hit the same symmetry sector), stages
//                   operand tiles in shared memory and issues mma.sync.m8n8k4.f64 (DMMA).  tcgen05 has no
//                   FP64 kind, so DMMA is the FP64 tensor path of sm_100a.
//   k_wapply_dmma   the W application as grouped small dense products on DMMA: destination panels fed by the
//                   same source panels share one pass over those sources (SU2 Wigner-9j couplings and
//                   Hermitian phases are folded into the coefficients on the host)
//   k_vec_*         solver-side BLAS-1 on device-resident vectors
//
// There is no CPU fallback: every entry point fails with a status when the device is not usable.
#include "../../include/qcm_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

// --------------------------------------------------------------------------------------------------------
// error handling
static thread_local std::string g_err;
static int fail(std::string const& s) { g_err = s; return 1; }
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define CHECK_INIT() do { if (!G.ready) return fail("qcm_init has not been called (or failed): no usable CUDA device"); } while (0)

struct qcm_array_s { double* p; int64_t n; };

struct BufTable { double* p[QCM_BUF_COUNT]; };

// device-side task records ---------------------------------------------------------------------------------
struct DSeg { long long a_off, b_off; int a_buf, b_buf, lda, ldb, m, n, k, ta, tb, pad; double alpha; };
struct DWork { long long c_off; int c_buf, ldc, m0, n0, m, n, seg_begin, seg_end, mode, pad; };   // mode 0 store, 1 add, 2 atomic
struct DWSrc { long long off; int buf, lds; };
struct DWDst { long long off; int buf, ldd; };
struct DWGroup { int rows, cols, n_src, n_dst, ng, src_begin, dst_begin, tpc; long long coef_begin; };   // tpc = 8-row tiles per column
struct DWWork { int group, r8, c0, pad; };
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

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// W application.  One WARP owns one work item: a strip of WT consecutive panel columns x 8 consecutive rows of a
// group, and streams over the group's source panels four at a time.  DMMA shape: M = 8 panel rows, K = 4 sources,
// N = 8 destinations (two N tiles in the WIDE variant for groups of 9..16 destinations).  A fragments come
// straight from global memory (every source element is needed exactly once per group), B fragments are the
// coefficients.  The loop is software pipelined: source descriptors are fetched two chunks ahead, source elements
// one chunk ahead of the DMMAs that consume them.
constexpr int W_WARPS = 4;
template <bool WIDE>
__global__ void __launch_bounds__(W_WARPS * 32, 5)
k_wapply_dmma(const DWWork* __restrict__ works, int n_works, const DWGroup* __restrict__ groups, const DWSrc* __restrict__ srcs,
              const DWDst* __restrict__ dsts, const double* __restrict__ coefs, const __grid_constant__ BufTable bufs)
{
    constexpr int WT = WIDE ? 4 : 8;
    constexpr int NG = WIDE ? 16 : 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wi = blockIdx.x * W_WARPS + warp;
    if (wi >= n_works) return;
    const DWWork w = works[wi];
    const DWGroup g = groups[w.group];
    const int fr = lane >> 2, fk = lane & 3;
    const int r = w.r8 * 8 + fr;
    const bool rok = r < g.rows;
    const int ncv = min(WT, g.cols - w.c0);
    const DWSrc* __restrict__ sp = srcs + g.src_begin;
    const double* __restrict__ cp = coefs + g.coef_begin + fr;

    double acc[WT][WIDE ? 2 : 1][2];
#pragma unroll
    for (int i = 0; i < WT; ++i)
#pragma unroll
        for (int j = 0; j < (WIDE ? 2 : 1); ++j) acc[i][j][0] = acc[i][j][1] = 0.;

    // pipeline registers
    const double* p_nn = nullptr; int lds_nn = 0;      // descriptor of chunk (cur + 2)
    const double* p_n = nullptr; int lds_n = 0;        // descriptor of chunk (cur + 1)
    double a_n[WT];                                     // elements of chunk (cur + 1)
    double b_n[WIDE ? 2 : 1];
    auto fetch_desc = [&](int u0, const double*& p, int& lds) {
        const int u = u0 + fk;
        p = nullptr; lds = 0;
        if (u < g.n_src && rok) { DWSrc q = sp[u]; p = bufs.p[q.buf] + q.off + r + (long long)w.c0 * q.lds; lds = q.lds; }
    };
    auto fetch_elems = [&](int u0, const double* p, int lds, double (&a)[WT], double (&b)[WIDE ? 2 : 1]) {
#pragma unroll
        for (int i = 0; i < WT; ++i) a[i] = (p != nullptr && i < ncv) ? p[(long long)i * lds] : 0.;
        const int u = u0 + fk;                         // coefficient rows are padded to a multiple of 4 sources
        b[0] = u0 < g.n_src ? cp[(long long)u * NG] : 0.;
        if (WIDE) b[WIDE ? 1 : 0] = u0 < g.n_src ? cp[(long long)u * NG + 8] : 0.;
    };
    fetch_desc(0, p_n, lds_n);
    fetch_desc(4, p_nn, lds_nn);
    fetch_elems(0, p_n, lds_n, a_n, b_n);
    for (int u0 = 0; u0 < g.n_src; u0 += 4) {
        double a[WT], b[WIDE ? 2 : 1];
#pragma unroll
        for (int i = 0; i < WT; ++i) a[i] = a_n[i];
        b[0] = b_n[0];
        if (WIDE) b[WIDE ? 1 : 0] = b_n[WIDE ? 1 : 0];
        p_n = p_nn; lds_n = lds_nn;
        fetch_desc(u0 + 8, p_nn, lds_nn);
        fetch_elems(u0 + 4, p_n, lds_n, a_n, b_n);
#pragma unroll
        for (int i = 0; i < WT; ++i) {
            dmma8x8x4(acc[i][0][0], acc[i][0][1], a[i], b[0]);
            if (WIDE) dmma8x8x4(acc[i][WIDE ? 1 : 0][0], acc[i][WIDE ? 1 : 0][1], a[i], b[WIDE ? 1 : 0]);
        }
    }
    // C fragment: row = panel row (lane/4), cols = destinations 2*(lane%4) + {0,1} (+8 for the second N tile)
    if (!rok) return;
#pragma unroll
    for (int j = 0; j < (WIDE ? 2 : 1); ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int d = j * 8 + 2 * fk + e;
            if (d >= g.n_dst) continue;
            const DWDst q = dsts[g.dst_begin + d];
            double* __restrict__ o = bufs.p[q.buf] + q.off + r + (long long)w.c0 * q.ldd;
#pragma unroll
            for (int i = 0; i < WT; ++i)
                if (i < ncv) o[(long long)i * q.ldd] = acc[i][j][e];
        }
}

constexpr int KC = 16;         // K chunk staged per pipeline stage
constexpr int SPAD = 4;        // row padding: (TM + 4) % 16 == 4 makes the fragment reads conflict free
constexpr int STAGES = 3;

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = valid ? 8 : 0;    // src-size 0: nothing is read, the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Grouped, variable-size FP64 GEMM.  CTA = WARPS_M x WARPS_N warps, each warp owns WMT x WNT DMMA tiles (8x8).
// One CTA computes one output tile and walks the K-segments of its work item (the terms of the sum over the MPO
// bond index that land in this symmetry sector).  Operand tiles are staged in shared memory by a STAGES-deep
// cp.async pipeline that runs across segment borders; alpha (Hermitian phase / conjugate correction) is applied
// to the A fragments.
constexpr int LDK = KC + 4;     // K-major rows: (KC + 4) % 16 == 4, conflict free as well
template <int WARPS_M, int WARPS_N, int WMT, int WNT>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32)
k_gemm_dmma(const DWork* __restrict__ works, const DSeg* __restrict__ segs, const __grid_constant__ BufTable bufs)
{
    constexpr int TM = WARPS_M * WMT * 8, TN = WARPS_N * WNT * 8, NT = WARPS_M * WARPS_N * 32;
    constexpr int LDA_S = TM + SPAD, LDB_S = TN + SPAD;
    // a staged operand tile is stored in the orientation in which global memory is contiguous:
    //   "M-major"  [kk][mm]  when the m (resp. n) index is contiguous in memory (A not transposed, B transposed)
    //   "K-major"  [mm][kk]  when the k index is contiguous (A transposed, B not transposed)
    constexpr int A_STAGE = (KC * LDA_S > TM * LDK) ? KC * LDA_S : TM * LDK;
    constexpr int B_STAGE = (KC * LDB_S > TN * LDK) ? KC * LDB_S : TN * LDK;
    extern __shared__ double smem[];
    double* As = smem;                          // [STAGES][A_STAGE]
    double* Bs = smem + STAGES * A_STAGE;       // [STAGES][B_STAGE]
    __shared__ double alpha_s[STAGES];
    __shared__ int flags_s[STAGES];             // bit 0: A is K-major, bit 1: B is K-major

    const DWork w = works[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int fr = lane >> 2, fk = lane & 3;

    double acc[WMT][WNT][2];
#pragma unroll
    for (int i = 0; i < WMT; ++i)
#pragma unroll
        for (int j = 0; j < WNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.;

    // producer cursor over the flattened (segment, k-chunk) sequence
    int ps = w.seg_begin, pk0 = 0;
    int p_lda = 0, p_ldb = 0, p_k = 0, p_ta = 0, p_tb = 0, pmrem = 0, pnrem = 0;
    double p_alpha = 0.;
    const double* __restrict__ pA = nullptr;
    const double* __restrict__ pB = nullptr;
    int nchunks = 0;
    auto load_seg = [&]() {
        while (ps < w.seg_end) {
            const DSeg sg = segs[ps];
            pmrem = sg.m - w.m0; pnrem = sg.n - w.n0;
            if (pmrem > 0 && pnrem > 0 && sg.k > 0) {          // segments may be smaller than the output block
                p_lda = sg.lda; p_ldb = sg.ldb; p_k = sg.k; p_ta = sg.ta; p_tb = sg.tb; p_alpha = sg.alpha;
                // tile origin folded into the base pointers
                pA = bufs.p[sg.a_buf] + sg.a_off + (sg.ta ? (long long)w.m0 * sg.lda : (long long)w.m0);
                pB = bufs.p[sg.b_buf] + sg.b_off + (sg.tb ? (long long)w.n0 : (long long)w.n0 * sg.ldb);
                pk0 = 0;
                return;
            }
            ++ps;
        }
    };
    load_seg();
    auto issue = [&](int stage) {
        if (ps < w.seg_end) {
            double* as = As + stage * A_STAGE;
            double* bs = Bs + stage * B_STAGE;
            const int krem = p_k - pk0;
            if (!p_ta) {
                const double* __restrict__ base = pA + (long long)pk0 * p_lda;
#pragma unroll
                for (int idx = tid; idx < TM * KC; idx += NT) {
                    const int mm = idx % TM, kk = idx / TM;
                    const bool v = mm < pmrem && kk < krem;
                    cp_async8(as + kk * LDA_S + mm, v ? base + (mm + kk * p_lda) : pA, v);
                }
            } else {
                const double* __restrict__ base = pA + pk0;
#pragma unroll
                for (int idx = tid; idx < TM * KC; idx += NT) {
                    const int kk = idx % KC, mm = idx / KC;
                    const bool v = mm < pmrem && kk < krem;
                    cp_async8(as + mm * LDK + kk, v ? base + (kk + mm * p_lda) : pA, v);
                }
            }
            if (!p_tb) {
                const double* __restrict__ base = pB + pk0;
#pragma unroll
                for (int idx = tid; idx < TN * KC; idx += NT) {
                    const int kk = idx % KC, nn = idx / KC;
                    const bool v = nn < pnrem && kk < krem;
                    cp_async8(bs + nn * LDK + kk, v ? base + (kk + nn * p_ldb) : pB, v);
                }
            } else {
                const double* __restrict__ base = pB + (long long)pk0 * p_ldb;
#pragma unroll
                for (int idx = tid; idx < TN * KC; idx += NT) {
                    const int nn = idx % TN, kk = idx / TN;
                    const bool v = nn < pnrem && kk < krem;
                    cp_async8(bs + kk * LDB_S + nn, v ? base + (nn + kk * p_ldb) : pB, v);
                }
            }
            if (tid == 0) { alpha_s[stage] = p_alpha; flags_s[stage] = (p_ta ? 1 : 0) | (p_tb ? 0 : 2); }
            pk0 += KC;
            if (pk0 >= p_k) { ++ps; load_seg(); }
        }
        cp_async_commit();
    };

    // total number of chunks this work item will consume (same walk as the producer, without loading)
    for (int s = w.seg_begin; s < w.seg_end; ++s) {
        const int m = segs[s].m, n = segs[s].n, k = segs[s].k;
        if (m - w.m0 > 0 && n - w.n0 > 0 && k > 0) nchunks += (k + KC - 1) / KC;
    }

#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st) issue(st);
    for (int it = 0; it < nchunks; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue((it + STAGES - 1) % STAGES);      // refills the stage consumed in the previous iteration
        const int stage = it % STAGES;
        const double* as = As + stage * A_STAGE;
        const double* bs = Bs + stage * B_STAGE;
        const double alpha = alpha_s[stage];
        const int fl = flags_s[stage];
        // per-lane fragment base and strides for the two orientations
        const int a_base = (fl & 1) ? (wm * WMT * 8 + fr) * LDK + fk : fk * LDA_S + wm * WMT * 8 + fr;
        const int a_ti = (fl & 1) ? 8 * LDK : 8, a_tk = (fl & 1) ? 4 : 4 * LDA_S;
        const int b_base = (fl & 2) ? (wn * WNT * 8 + fr) * LDK + fk : fk * LDB_S + wn * WNT * 8 + fr;
        const int b_tj = (fl & 2) ? 8 * LDK : 8, b_tk = (fl & 2) ? 4 : 4 * LDB_S;
#pragma unroll
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            double a[WMT], b[WNT];
#pragma unroll
            for (int i = 0; i < WMT; ++i) a[i] = alpha * as[a_base + i * a_ti + k4 * a_tk];
#pragma unroll
            for (int j = 0; j < WNT; ++j) b[j] = bs[b_base + j * b_tj + k4 * b_tk];
#pragma unroll
            for (int i = 0; i < WMT; ++i)
#pragma unroll
                for (int j = 0; j < WNT; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    // ---- epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1})
    double* __restrict__ C = bufs.p[w.c_buf] + w.c_off;
#pragma unroll
    for (int i = 0; i < WMT; ++i)
#pragma unroll
        for (int j = 0; j < WNT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = w.m0 + (wm * WMT + i) * 8 + fr, c = w.n0 + (wn * WNT + j) * 8 + 2 * fk + e;
                if (r < w.m && c < w.n) {
                    double* q = C + r + (long long)c * w.ldc;
                    if (w.mode == 0) *q = acc[i][j][e];
                    else if (w.mode == 1) *q += acc[i][j][e];
                    else atomicAdd(q, acc[i][j][e]);
                }
            }
}

// solver-side BLAS-1 --------------------------------------------------------------------------------------
__global__ void k_vec_dot(const double* __restrict__ x, const double* __restrict__ y, long long n, double* out)
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
        if (threadIdx.x == 0) atomicAdd(out, s);
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
struct AxpyGroup { DWWork* d_works = nullptr; DWGroup* d_groups = nullptr; DWSrc* d_srcs = nullptr; DWDst* d_dsts = nullptr; double* d_coefs = nullptr; int64_t n_works = 0, n_narrow = 0; };
struct WaveDev { GemmGroup t, c; AxpyGroup w; int64_t y_elems = 0, t_elems = 0; };

struct qcm_plan_s
{
    int kind = 0;
    DCopy* d_copies = nullptr; int64_t n_copies = 0;
    GemmGroup p;
    std::vector<WaveDev> waves;
    int64_t elems[QCM_BUF_COUNT];
    double flops = 0; int64_t bytes = 0;
    int64_t n_launches = 0;
    std::vector<void*> allocs;
};

typedef int (*nccl_get_uid_t)(void*);
typedef int (*nccl_init_rank_t)(void**, int, char[128], int);   // ncclUniqueId is passed BY VALUE (128-byte struct)
struct NcclUid { char internal[128]; };
typedef int (*nccl_get_uid_fn)(NcclUid*);
typedef int (*nccl_init_rank_fn)(void**, int, NcclUid, int);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
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
    bool timing = false;
    double last_ms[6] = {0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8];
    // nccl (resolved at run time so that single-GPU use has no NCCL dependency)
    void* nccl_lib = nullptr; void* comm = nullptr; int rank = 0, world = 1;
    nccl_get_uid_fn f_uid = nullptr; nccl_init_rank_fn f_init = nullptr; nccl_allreduce_fn f_ar = nullptr;
    nccl_destroy_fn f_destroy = nullptr; nccl_errstr_fn f_err = nullptr;
} G;

static int gemm_set_attributes();
static int ensure_ws(int slot, int64_t n)
{
    if (n <= G.ws_elems[slot]) return 0;
    if (G.ws[slot]) { CU(cudaStreamSynchronize(G.stream)); CU(cudaFree(G.ws[slot])); G.ws[slot] = nullptr; G.ws_elems[slot] = 0; }
    int64_t want = n + n / 8 + 1024;
    cudaError_t e = cudaMalloc((void**)&G.ws[slot], (size_t)want * sizeof(double));
    if (e != cudaSuccess) {
        want = n;
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
    for (auto& ev : G.ev) CU(cudaEventCreate(&ev));
    for (int i = 0; i < Global::kAux; ++i) { CU(cudaStreamCreateWithFlags(&G.aux[i], cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&G.join_ev[i], cudaEventDisableTiming)); }
    CU(cudaEventCreateWithFlags(&G.fork_ev, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&G.scratch, 4096));
    if (gemm_set_attributes()) return 1;
    G.device = device;
    G.ready = true;
    return 0;
}

extern "C" int qcm_finalize(void)
{
    if (!G.ready) return 0;
    cudaStreamSynchronize(G.stream);
    for (int i = 0; i < QCM_BUF_COUNT; ++i) if (G.ws[i]) { cudaFree(G.ws[i]); G.ws[i] = nullptr; G.ws_elems[i] = 0; }
    if (G.scratch) cudaFree(G.scratch);
    G.scratch = nullptr;
    for (auto& ev : G.ev) cudaEventDestroy(ev);
    for (int i = 0; i < Global::kAux; ++i) { cudaStreamDestroy(G.aux[i]); cudaEventDestroy(G.join_ev[i]); }
    cudaEventDestroy(G.fork_ev);
    cudaStreamDestroy(G.stream);
    G.stream = nullptr; G.ready = false; G.device = -1;
    return 0;
}

extern "C" const char* qcm_last_error(void) { return g_err.c_str(); }
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
    qcm_array_s* a = new qcm_array_s{nullptr, n};
    if (n > 0) {
        cudaError_t e = cudaMalloc((void**)&a->p, (size_t)n * sizeof(double));
        if (e != cudaSuccess) { delete a; return fail(std::string("qcm_array_alloc: ") + cudaGetErrorString(e)); }
    }
    *out = a;
    return 0;
}
extern "C" int qcm_array_free(qcm_array_t a)
{
    if (!a) return 0;
    if (a->p) { cudaStreamSynchronize(G.stream); cudaFree(a->p); }
    delete a;
    return 0;
}
extern "C" int qcm_array_size(qcm_array_t a, int64_t* n) { if (!a) return fail("null array"); *n = a->n; return 0; }
extern "C" void* qcm_array_devptr(qcm_array_t a) { return a ? (void*)a->p : nullptr; }
extern "C" int qcm_array_upload(qcm_array_t a, int64_t off, const double* host, int64_t n)
{
    CHECK_INIT();
    if (!a || off < 0 || n < 0 || off + n > a->n) return fail("qcm_array_upload: range outside the array");
    if (n == 0) return 0;
    CU(cudaMemcpyAsync(a->p + off, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}
extern "C" int qcm_array_download(qcm_array_t a, int64_t off, double* host, int64_t n)
{
    CHECK_INIT();
    if (!a || off < 0 || n < 0 || off + n > a->n) return fail("qcm_array_download: range outside the array");
    if (n == 0) return 0;
    CU(cudaMemcpyAsync(host, a->p + off, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}
extern "C" int qcm_array_zero(qcm_array_t a)
{
    CHECK_INIT();
    if (!a) return fail("null array");
    if (a->n) CU(cudaMemsetAsync(a->p, 0, (size_t)a->n * sizeof(double), G.stream));
    return 0;
}

// ---- plan construction ----------------------------------------------------------------------------------
struct TileVariant { int tm, tn, threads; double eff; };
static const TileVariant kVariants[] = {
    {64, 128, 256, 1.00},  // 0: 2x4 warps, 4x4 tiles
    {128, 64, 256, 1.00},  // 1: 4x2 warps, 4x4
    {64, 64, 128, 0.90},   // 2: 2x2 warps, 4x4
    {32, 128, 128, 0.85},  // 3: 1x4 warps, 4x4
    {128, 32, 128, 0.85},  // 4: 4x1 warps, 4x4
    {16, 128, 128, 0.60},  // 5: 1x4 warps, 2x4
    {128, 16, 128, 0.60},  // 6: 4x1 warps, 4x2
    {32, 32, 128, 0.45},   // 7: 2x2 warps, 2x2
    {16, 16, 128, 0.20},   // 8: 2x2 warps, 1x1
    {8, 128, 128, 0.35},   // 9: 1x4 warps, 1x4
    {128, 8, 128, 0.35},   // 10: 4x1 warps, 4x1
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
static size_t variant_smem(int v)
{
    size_t a = std::max(KC * (kVariants[v].tm + SPAD), kVariants[v].tm * LDK), b = std::max(KC * (kVariants[v].tn + SPAD), kVariants[v].tn * LDK);
    return (size_t)STAGES * (a + b) * sizeof(double);
}

#define QCM_FOR_EACH_VARIANT(X) \
    X(0, 2, 4, 4, 4) X(1, 4, 2, 4, 4) X(2, 2, 2, 4, 4) X(3, 1, 4, 4, 4) X(4, 4, 1, 4, 4) X(5, 1, 4, 2, 4) X(6, 4, 1, 4, 2) \
    X(7, 2, 2, 2, 2) X(8, 2, 2, 1, 1) X(9, 1, 4, 1, 4) X(10, 4, 1, 4, 1)

static int gemm_set_attributes()
{
#define X(v, a, b, c, d) CU(cudaFuncSetAttribute(k_gemm_dmma<a, b, c, d>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)variant_smem(v)));
    QCM_FOR_EACH_VARIANT(X)
#undef X
    return 0;
}

static void launch_gemm_variant(int v, int64_t n, const DWork* works, const DSeg* segs, BufTable const& bufs, cudaStream_t st)
{
    dim3 g((unsigned)n), b(kVariants[v].threads);
    size_t sm = variant_smem(v);
    switch (v) {
#define X(vv, a, bb, c, d) case vv: k_gemm_dmma<a, bb, c, d><<<g, b, sm, st>>>(works, segs, bufs); break;
    QCM_FOR_EACH_VARIANT(X)
#undef X
    }
}

static int pick_variant(int m, int n)
{
    int best = 0; double best_cost = 1e300;
    for (int v = 0; v < kNumVariants; ++v) {
        double tiles = (double)((m + kVariants[v].tm - 1) / kVariants[v].tm) * (double)((n + kVariants[v].tn - 1) / kVariants[v].tn);
        double cost = tiles * kVariants[v].tm * kVariants[v].tn / kVariants[v].eff;
        if (cost < best_cost) { best_cost = cost; best = v; }
    }
    return best;
}

template <class T> static int dev_upload(qcm_plan_s* P, std::vector<T> const& h, T** d)
{
    *d = nullptr;
    if (h.empty()) return 0;
    CU(cudaMalloc((void**)d, h.size() * sizeof(T)));
    P->allocs.push_back(*d);
    CU(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// output blocks -> tile work items; long segment lists are split (split-K over the MPO bond index) and
// combined with FP64 atomics
static int build_gemm_group(qcm_plan_s* P, GemmGroup& g, const qcm_gemm_out* outs, int64_t n_outs, const qcm_gemm_seg* segs, int64_t n_segs,
                            int base_mode /*0 store, 1 add*/)
{
    std::vector<DSeg> hs((size_t)n_segs);
    for (int64_t i = 0; i < n_segs; ++i) {
        qcm_gemm_seg const& s = segs[i];
        hs[i] = DSeg{s.A.off, s.B.off, s.A.buf, s.B.buf, s.lda, s.ldb, s.m, s.n, s.k, s.ta, s.tb, 0, s.alpha};
    }
    std::vector<std::vector<DWork>> per_variant(kNumVariants);
    const int max_chunks = 96;       // K chunks (of KC) per work item before the segment list is split
    for (int64_t o = 0; o < n_outs; ++o) {
        qcm_gemm_out const& out = outs[o];
        if (out.m <= 0 || out.n <= 0) continue;
        int v = pick_variant(out.m, out.n);
        // chunk the segment list
        std::vector<std::pair<int, int>> chunks;
        {
            int csum = 0, cb = out.seg_begin;
            for (int s = out.seg_begin; s < out.seg_end; ++s) {
                csum += (segs[s].k + KC - 1) / KC;
                if (csum >= max_chunks && s + 1 < out.seg_end) { chunks.push_back(std::make_pair(cb, s + 1)); cb = s + 1; csum = 0; }
            }
            chunks.push_back(std::make_pair(cb, out.seg_end));
        }
        int mode = chunks.size() > 1 ? 2 : base_mode;
        if (chunks.size() > 1 && base_mode == 0) return fail("internal: split-K on a store-mode output");
        for (auto const& ch : chunks)
            for (int n0 = 0; n0 < out.n; n0 += kVariants[v].tn)
                for (int m0 = 0; m0 < out.m; m0 += kVariants[v].tm)
                    per_variant[v].push_back(DWork{out.C.off, out.C.buf, out.ldc, m0, n0, out.m, out.n, ch.first, ch.second, mode, 0});
    }
    std::vector<DWork> hw;
    for (int v = 0; v < kNumVariants; ++v) {
        if (per_variant[v].empty()) continue;
        // heavy work first: better tail behaviour of the hardware scheduler
        std::stable_sort(per_variant[v].begin(), per_variant[v].end(), [&](DWork const& a, DWork const& b) {
            return (a.seg_end - a.seg_begin) > (b.seg_end - b.seg_begin);
        });
        g.launches.push_back(GemmLaunch{v, (int64_t)hw.size(), (int64_t)per_variant[v].size()});
        hw.insert(hw.end(), per_variant[v].begin(), per_variant[v].end());
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
    for (int64_t i = 0; i < wd.n_w_srcs; ++i) hs[i] = DWSrc{wd.w_srcs[i].src.off, wd.w_srcs[i].src.buf, wd.w_srcs[i].lds};
    std::vector<DWDst> hd((size_t)wd.n_w_dsts);
    for (int64_t i = 0; i < wd.n_w_dsts; ++i) hd[i] = DWDst{wd.w_dsts[i].dst.off, wd.w_dsts[i].dst.buf, wd.w_dsts[i].ldd};
    std::vector<DWGroup> hg((size_t)wd.n_w_groups);
    std::vector<DWWork> hw, hw_wide;      // one work item per warp: 8 rows x WT columns of the group's panels
    for (int64_t i = 0; i < wd.n_w_groups; ++i) {
        qcm_w_group const& q = wd.w_groups[i];
        if (q.ng != 8 && q.ng != 16) return fail("qcm_plan_create: W group with ng outside {8,16}");
        if (q.n_dst > q.ng) return fail("qcm_plan_create: W group with more destinations than ng");
        int tpc = (q.rows + 7) / 8;
        hg[i] = DWGroup{q.rows, q.cols, q.n_src, q.n_dst, q.ng, q.src_begin, q.dst_begin, tpc, q.coef_begin};
        const int wt = q.ng == 16 ? 4 : 8;
        std::vector<DWWork>& dstv = q.ng == 16 ? hw_wide : hw;
        for (int c0 = 0; c0 < q.cols; c0 += wt)
            for (int r8 = 0; r8 < tpc; ++r8) dstv.push_back(DWWork{(int)i, r8, c0, 0});
    }
    g.n_narrow = (int64_t)hw.size();
    hw.insert(hw.end(), hw_wide.begin(), hw_wide.end());
    std::vector<double> hc(wd.w_coefs, wd.w_coefs + wd.n_w_coefs);
    g.n_works = (int64_t)hw.size();
    if (dev_upload(P, hw, &g.d_works) || dev_upload(P, hg, &g.d_groups) || dev_upload(P, hs, &g.d_srcs) || dev_upload(P, hd, &g.d_dsts) ||
        dev_upload(P, hc, &g.d_coefs)) return 1;
    if (g.n_narrow) P->n_launches += 1;
    if (g.n_works > g.n_narrow) P->n_launches += 1;
    return 0;
}

extern "C" int qcm_plan_create(const qcm_plan_desc* d, qcm_plan_t* out)
{
    CHECK_INIT();
    if (!d || !out) return fail("qcm_plan_create: null argument");
    qcm_plan_s* P = new qcm_plan_s();
    P->kind = d->kind; P->flops = d->flops; P->bytes = d->bytes;
    for (int i = 0; i < QCM_BUF_COUNT; ++i) P->elems[i] = d->elems[i];
    auto bail = [&]() { for (void* p : P->allocs) cudaFree(p); delete P; return 1; };
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
        W.y_elems = wd.y_elems; W.t_elems = wd.t_elems;
        if (build_gemm_group(P, W.t, wd.t_outs, wd.n_t_outs, wd.t_segs, wd.n_t_segs, 0)) return bail();
        if (build_axpy_group(P, W.w, wd)) return bail();
        if (build_gemm_group(P, W.c, wd.c_outs, wd.n_c_outs, wd.c_segs, wd.n_c_segs, 1)) return bail();
    }
    *out = P;
    return 0;
}

extern "C" int qcm_plan_destroy(qcm_plan_t P)
{
    if (!P) return 0;
    if (G.ready) cudaStreamSynchronize(G.stream);
    for (void* p : P->allocs) cudaFree(p);
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
        launch_gemm_variant(l.variant, l.n_works, g.d_works + l.work_begin, g.d_segs, bufs, st);
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
    float acc[4] = {0, 0, 0, 0};
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
    for (auto const& W : P->waves) {
        mark(3);
        if (run_gemm_group(W.t, bufs)) return 1;
        mark(4); lap(1, 3, 4);
        if (W.y_elems) CU(cudaMemsetAsync(bufs.p[QCM_BUF_Y], 0, (size_t)W.y_elems * 8, G.stream));
        if (W.w.n_narrow) {
            k_wapply_dmma<false><<<(unsigned)((W.w.n_narrow + W_WARPS - 1) / W_WARPS), W_WARPS * 32, 0, G.stream>>>(W.w.d_works, (int)W.w.n_narrow, W.w.d_groups, W.w.d_srcs, W.w.d_dsts, W.w.d_coefs, bufs);
            G.launches++;
        }
        if (W.w.n_works > W.w.n_narrow) {
            int64_t nw = W.w.n_works - W.w.n_narrow;
            k_wapply_dmma<true><<<(unsigned)((nw + W_WARPS - 1) / W_WARPS), W_WARPS * 32, 0, G.stream>>>(W.w.d_works + W.w.n_narrow, (int)nw, W.w.d_groups, W.w.d_srcs, W.w.d_dsts, W.w.d_coefs, bufs);
            G.launches++;
        }
        mark(5); lap(2, 4, 5);
        if (run_gemm_group(W.c, bufs)) return 1;
        mark(6); lap(3, 5, 6);
    }
    CU(cudaGetLastError());
    if (tm) { for (int i = 0; i < 4; ++i) G.last_ms[i] = acc[i]; G.last_ms[4] = 0; G.last_ms[5] = acc[0] + acc[1] + acc[2] + acc[3]; }
    return 0;
}

static int check_arr(qcm_array_t a, int64_t need, const char* what)
{
    if (!a) return fail(std::string(what) + ": null array");
    if (a->n < need) return fail(std::string(what) + ": array holds " + std::to_string(a->n) + " elements, plan needs " + std::to_string(need));
    return 0;
}

static int allreduce_ptr(double* p, int64_t n);

extern "C" int qcm_site_hamil2_dev(qcm_plan_t P, qcm_array_t left, qcm_array_t right, qcm_array_t psi, qcm_array_t sigma)
{
    CHECK_INIT();
    if (!P || P->kind != 0) return fail("qcm_site_hamil2: plan is not a sigma plan");
    if (check_arr(left, P->elems[QCM_BUF_LEFT], "left boundary") || check_arr(right, P->elems[QCM_BUF_RIGHT], "right boundary") ||
        check_arr(psi, P->elems[QCM_BUF_KET_LP], "psi") || check_arr(sigma, P->elems[QCM_BUF_OUT], "sigma")) return 1;
    BufTable b; memset(&b, 0, sizeof(b));
    b.p[QCM_BUF_LEFT] = left->p; b.p[QCM_BUF_RIGHT] = right->p; b.p[QCM_BUF_KET_LP] = psi->p; b.p[QCM_BUF_OUT] = sigma->p;
    if (P->elems[QCM_BUF_OUT]) CU(cudaMemsetAsync(sigma->p, 0, (size_t)P->elems[QCM_BUF_OUT] * 8, G.stream));
    if (execute(P, b)) return 1;
    if (G.world > 1) return allreduce_ptr(sigma->p, P->elems[QCM_BUF_OUT]);
    return 0;
}

extern "C" int qcm_site_hamil2(qcm_plan_t P, qcm_array_t left, qcm_array_t right, const double* psi, double* sigma)
{
    CHECK_INIT();
    if (!P || P->kind != 0) return fail("qcm_site_hamil2: plan is not a sigma plan");
    if (ensure_ws(QCM_BUF_KET_LP, P->elems[QCM_BUF_KET_LP]) || ensure_ws(QCM_BUF_OUT, P->elems[QCM_BUF_OUT])) return 1;
    qcm_array_s a_psi{G.ws[QCM_BUF_KET_LP], G.ws_elems[QCM_BUF_KET_LP]}, a_sig{G.ws[QCM_BUF_OUT], G.ws_elems[QCM_BUF_OUT]};
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
    if (check_arr(in, P->elems[in_slot], "input boundary") || check_arr(out, P->elems[QCM_BUF_OUT], "output boundary")) return 1;
    if (ensure_ws(QCM_BUF_KET_LP, P->elems[QCM_BUF_KET_LP]) || ensure_ws(QCM_BUF_BRA_LP, P->elems[QCM_BUF_BRA_LP])) return 1;
    if (P->elems[QCM_BUF_KET_LP]) CU(cudaMemcpyAsync(G.ws[QCM_BUF_KET_LP], ket, (size_t)P->elems[QCM_BUF_KET_LP] * 8, cudaMemcpyHostToDevice, G.stream));
    if (P->elems[QCM_BUF_BRA_LP]) CU(cudaMemcpyAsync(G.ws[QCM_BUF_BRA_LP], bra, (size_t)P->elems[QCM_BUF_BRA_LP] * 8, cudaMemcpyHostToDevice, G.stream));
    BufTable b; memset(&b, 0, sizeof(b));
    b.p[in_slot] = in->p; b.p[QCM_BUF_KET_LP] = G.ws[QCM_BUF_KET_LP]; b.p[QCM_BUF_BRA_LP] = G.ws[QCM_BUF_BRA_LP]; b.p[QCM_BUF_OUT] = out->p;
    if (out->n) CU(cudaMemsetAsync(out->p, 0, (size_t)out->n * 8, G.stream));
    if (execute(P, b)) return 1;
    // every rank computed its share of the output bond indices into a zeroed array: the sum is the full boundary
    if (G.world > 1 && allreduce_ptr(out->p, P->elems[QCM_BUF_OUT])) return 1;
    CU(cudaStreamSynchronize(G.stream));
    return 0;
}

extern "C" int qcm_set_timing(int enabled) { G.timing = enabled != 0; return 0; }
extern "C" int qcm_last_timing(double ms[6]) { for (int i = 0; i < 6; ++i) ms[i] = G.last_ms[i]; return 0; }

// ---- BLAS-1 ---------------------------------------------------------------------------------------------
extern "C" int qcm_vec_dot(qcm_array_t x, qcm_array_t y, int64_t n, double* result)
{
    CHECK_INIT();
    if (check_arr(x, n, "x") || check_arr(y, n, "y")) return 1;
    CU(cudaMemsetAsync(G.scratch, 0, 8, G.stream));
    if (n) { k_vec_dot<<<std::max(1, std::min(G.sm_count * 4, (int)((n + 255) / 256))), 256, 0, G.stream>>>(x->p, y->p, n, G.scratch); G.launches++; }
    CU(cudaMemcpyAsync(result, G.scratch, 8, cudaMemcpyDeviceToHost, G.stream));
    CU(cudaStreamSynchronize(G.stream));
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
    G.f_destroy = (nccl_destroy_fn)dlsym(h, "ncclCommDestroy");
    G.f_err = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
    if (!G.f_uid || !G.f_init || !G.f_ar || !G.f_destroy) { G.f_ar = nullptr; return fail("libnccl lacks required symbols"); }
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
        float ms = 0; cudaEventElapsedTime(&ms, G.ev[6], G.ev[7]); G.last_ms[4] = ms; G.last_ms[5] += ms;
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
