// wgemm_ws.cu -- the W application for high fan-in destination panels as gathered dense products on DMMA.
//
// Replaces the axpy panels of the reference's MPO-tensor kernels (lb_tensor_mpo / rb_tensor_mpo, alps_detail.hpp:189-224;
// SU2 detail::lbtm / rbtm / task_axpy, non-abelian/micro_kernels.hpp:19-198) where a destination panel of Y is a sum
// over many source panels of T (the integral-weighted sums over O(l^2) MPO bond terms).  Destinations fed by nearly
// the same sources form a group (host: Planner::group_axpy); a group is one product
//     dst[e, d] = sum_u src_u[e] * coef[u][d]        e: panel elements, u: sources, d: up to NG destinations
// with coef = W entry * term scale * Wigner-9j coupling * Hermitian phase (zero where a destination skips a source),
// so every source panel is read once per group instead of once per destination.
//
// Same execution scheme as gemm_ws.cu: persistent CTAs take (group, 128-element tile) work items round-robin; four
// producer warps take stages in turn and gather 16 sources per stage (each source row a coalesced run of 128 doubles)
// and the matching coefficient rows with cp.async into a shared-memory ring, signalled per stage through mbarriers
// (the low fan-out classes are HBM bound: four independent descriptor -> address -> copy chains keep the memory
// system busy); eight consumer warps
// issue mma.sync.m8n8k4.f64 (M = elements, N = destinations, K = sources) and store every destination directly from the
// accumulator fragments (eight consecutive elements per destination and quarter-warp).
#include "qcm_dev.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int KC = 16, SPAD = 4, NP = 4;

struct __align__(16) WMeta { int flags, group, e0, n, n_dst, pad0, pad1, pad2; };   // flags: bit 2 first, bit 3 last, bits 8..: depth; group -1: end

template <int WARPS_E, int WARPS_D, int WMT, int WNT, int STAGES>
struct WCfg
{
    static constexpr int TE = WARPS_E * WMT * 8, NG = WARPS_D * WNT * 8, NW = WARPS_E * WARPS_D, NT = (NW + NP) * 32;
    static constexpr int LDA = TE + SPAD, LDB = NG + SPAD;
    static constexpr int A_STAGE = KC * LDA, B_STAGE = KC * LDB;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + STAGES * (sizeof(WMeta) + 16);
};

template <int WARPS_E, int WARPS_D, int WMT, int WNT, int STAGES>
__global__ void __launch_bounds__((WARPS_E * WARPS_D + NP) * 32, 1)
k_wgemm_ws(const DWWork* __restrict__ works, int n_works, const DWGroup* __restrict__ groups, const DWSrc* __restrict__ srcs,
           const DWDst* __restrict__ dsts, const double* __restrict__ coefs, const __grid_constant__ BufTable bufs)
{
    using Cfg = WCfg<WARPS_E, WARPS_D, WMT, WNT, STAGES>;
    constexpr int TE = Cfg::TE, NG = Cfg::NG, NW = Cfg::NW, LDA = Cfg::LDA, LDB = Cfg::LDB, A_STAGE = Cfg::A_STAGE, B_STAGE = Cfg::B_STAGE;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                               // [STAGES][KC][LDA]
    double* Bs = smem + STAGES * A_STAGE;            // [STAGES][KC][LDB]
    WMeta* metas = reinterpret_cast<WMeta*>(smem + STAGES * (A_STAGE + B_STAGE));
    unsigned long long* full = reinterpret_cast<unsigned long long*>(metas + STAGES);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 32 + 1); mbar_init(&empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= NW) {
        // ------------------------------------------------------------------ producer warps
        // The producer warps take whole stages in turn (warp pw fills chunk numbers pw, pw + NP, ...): the dependent
        // chain "source descriptors -> addresses -> cp.async" of one stage overlaps with the chains of the next three.
        const int pw = warp - NW;
        long long cnt = 0;         // chunk number in the CTA's stream; stage = cnt % STAGES, phase = (cnt / STAGES) & 1
        for (int wi = blockIdx.x; wi < n_works; wi += gridDim.x) {
            const DWWork w = works[wi];
            const DWGroup g = groups[w.group];
            const int n = g.rows * g.cols;
            const int nchunks = (g.n_src + KC - 1) / KC;
            // chunks of this work item that fall to this warp
            int c0 = (int)((pw - cnt % NP + NP) % NP);
            if (c0 < nchunks) {
                // this lane's TE / 32 panel elements as (row, column) of the panel -- element offset inside a source is
                // row + column * lds (= the flat index when the panel is stored with lds == rows) -- and their byte counts
                int er[TE / 32], ec[TE / 32], nb[TE / 32];
#pragma unroll
                for (int j = 0; j < TE / 32; ++j) {
                    const int el = w.e0 + lane + 32 * j;
                    ec[j] = el / g.rows; er[j] = el - ec[j] * g.rows; nb[j] = el < n ? 8 : 0;
                }
                const DWSrc* __restrict__ sq = srcs + g.src_begin;
                const double* __restrict__ cq = coefs + g.coef_begin;
                for (int c = c0; c < nchunks; c += NP) {
                    const int u0 = c * KC;
                    const int krem = g.n_src - u0;
                    const int kld = min(KC, (krem + 3) & ~3);
                    const long long my = cnt + c;
                    const int stage = (int)(my % STAGES);
                    const unsigned phase = (unsigned)((my / STAGES) & 1);
                    const unsigned as = smem_u32(As + stage * A_STAGE) + lane * 8;
                    const unsigned bs = smem_u32(Bs + stage * B_STAGE);
                    bool waited = false;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {           // two halves of eight source rows: eight descriptor loads in flight
                        DWSrc q[KC / 2];
#pragma unroll
                        for (int t = 0; t < KC / 2; ++t) q[t] = sq[min(u0 + h * (KC / 2) + t, g.n_src - 1)];
                        if (!waited) {
                            mbar_wait(&empty[stage], phase ^ 1); waited = true;
                            // coefficient rows of the stage: KC * NG contiguous doubles (rows are zero-padded to a multiple of KC)
                            const double* __restrict__ cr = cq + (long long)u0 * NG;
#pragma unroll
                            for (int idx = lane; idx < KC * NG; idx += 32) {
                                const int kk = idx / NG, d = idx % NG;
                                if (kk < kld) cp_async8_s(bs + (kk * LDB + d) * 8, cr + idx, 8);
                            }
                        }
#pragma unroll
                        for (int t = 0; t < KC / 2; ++t) {
                            const int kk = h * (KC / 2) + t;
                            if (kk < kld) {
                                const bool kv = kk < krem;
                                const double* __restrict__ p = bufs.p[q[t].buf] + q[t].off;
#pragma unroll
                                for (int j = 0; j < TE / 32; ++j)
                                    cp_async8_s(as + (kk * LDA + 32 * j) * 8, p + (er[j] + ec[j] * q[t].lds), kv ? nb[j] : 0);
                            }
                        }
                    }
                    if (lane == 0) {
                        WMeta m; m.group = w.group; m.e0 = w.e0; m.n = n; m.n_dst = g.n_dst; m.pad0 = m.pad1 = m.pad2 = 0;
                        m.flags = (c == 0 ? 4 : 0) | (c == nchunks - 1 ? 8 : 0) | (kld << 8);
                        metas[stage] = m;
                    }
                    mbar_cp_async_arrive(&full[stage]);
                    if (lane == 0) mbar_arrive(&full[stage]);
                }
            }
            cnt += nchunks;
        }
        if (cnt % NP == pw) {       // end of stream
            const int stage = (int)(cnt % STAGES);
            const unsigned phase = (unsigned)((cnt / STAGES) & 1);
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) {
                WMeta m; m.flags = 0; m.group = -1; m.e0 = 0; m.n = 0; m.n_dst = 0; m.pad0 = m.pad1 = m.pad2 = 0;
                metas[stage] = m;
            }
            mbar_cp_async_arrive(&full[stage]);
            if (lane == 0) mbar_arrive(&full[stage]);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int we = warp % WARPS_E, wd = warp / WARPS_E;
    const int fr = lane >> 2, fk = lane & 3;
    const int row0 = we * WMT * 8, col0 = wd * WNT * 8;
    double acc[WMT][WNT][2];
    bool active = false;
    int stage = 0; unsigned phase = 0;
    for (;;) {
        mbar_wait(&full[stage], phase);
        const WMeta sm = metas[stage];
        if (sm.group < 0) break;
        const int fl = sm.flags;
        if (fl & 4) {
#pragma unroll
            for (int i = 0; i < WMT; ++i)
#pragma unroll
                for (int j = 0; j < WNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.;
            active = sm.e0 + row0 < sm.n && col0 < sm.n_dst;
        }
        const int k4n = fl >> 10;
        if (active) {
            const double* as = As + stage * A_STAGE + fk * LDA + row0 + fr;
            const double* bs = Bs + stage * B_STAGE + fk * LDB + col0 + fr;
            if (k4n == KC / 4) {
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    double a[WMT], b[WNT];
#pragma unroll
                    for (int i = 0; i < WMT; ++i) a[i] = as[i * 8 + k4 * 4 * LDA];
#pragma unroll
                    for (int j = 0; j < WNT; ++j) b[j] = bs[j * 8 + k4 * 4 * LDB];
#pragma unroll
                    for (int i = 0; i < WMT; ++i)
#pragma unroll
                        for (int j = 0; j < WNT; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            } else {
                for (int k4 = 0; k4 < k4n; ++k4) {
                    double a[WMT], b[WNT];
#pragma unroll
                    for (int i = 0; i < WMT; ++i) a[i] = as[i * 8 + k4 * 4 * LDA];
#pragma unroll
                    for (int j = 0; j < WNT; ++j) b[j] = bs[j * 8 + k4 * 4 * LDB];
#pragma unroll
                    for (int i = 0; i < WMT; ++i)
#pragma unroll
                        for (int j = 0; j < WNT; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if ((fl & 8) && active) {
            // ---- epilogue: fragment (element = lane/4, destinations = 2*(lane%4) + {0,1}); every destination panel has
            // its own base and leading dimension
            const DWGroup g = groups[sm.group];
            int eo[WMT], er[WMT], ec[WMT];
#pragma unroll
            for (int i = 0; i < WMT; ++i) {
                eo[i] = sm.e0 + row0 + i * 8 + fr;
                ec[i] = eo[i] / g.rows; er[i] = eo[i] - ec[i] * g.rows;
            }
#pragma unroll
            for (int j = 0; j < WNT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int d = col0 + j * 8 + 2 * fk + e;
                    if (d < sm.n_dst) {
                        const DWDst q = dsts[g.dst_begin + d];
                        double* __restrict__ p = bufs.p[q.buf] + q.off;
                        const bool flat = q.ldd == g.rows;
#pragma unroll
                        for (int i = 0; i < WMT; ++i)
                            if (eo[i] < sm.n) p[flat ? (long long)eo[i] : er[i] + (long long)ec[i] * q.ldd] = acc[i][j][e];
                    }
                }
        }
    }
}

//   X(class, WARPS_E, WARPS_D, WMT, WNT, STAGES)        128 panel elements x NG destinations
#define QCM_WG_VARIANTS(X) \
    X(0, 8, 1, 2, 1, 8) /* ng  8 */ \
    X(1, 4, 2, 4, 1, 8) /* ng 16 */ \
    X(2, 4, 2, 4, 2, 8) /* ng 32 */ \
    X(3, 4, 2, 4, 4, 8) /* ng 64 */

int g_sms = 0;
int g_occ[4];

}  // namespace

int wgemm_ws_tile() { return 128; }

const char* wgemm_ws_init(int sm_count)
{
    g_sms = sm_count;
    cudaError_t e;
#define X(c, a, b, wm, wn, s) \
    { using Cfg = WCfg<a, b, wm, wn, s>; \
      static_assert(Cfg::TE == 128, "work items are cut for 128 panel elements"); \
      e = cudaFuncSetAttribute(k_wgemm_ws<a, b, wm, wn, s>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM); \
      if (e != cudaSuccess) return cudaGetErrorString(e); \
      int occ = 0; \
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wgemm_ws<a, b, wm, wn, s>, Cfg::NT, Cfg::SMEM); \
      if (e != cudaSuccess) return cudaGetErrorString(e); \
      if (occ < 1) return "k_wgemm_ws variant does not fit on an SM"; \
      g_occ[c] = occ; \
      if (getenv("QCM_DEBUG")) fprintf(stderr, "wgemm_ws class %d: %d elements x %d destinations, %d threads, %zu B smem, %d CTAs/SM\n", c, Cfg::TE, Cfg::NG, Cfg::NT, (size_t)Cfg::SMEM, occ); }
    QCM_WG_VARIANTS(X)
#undef X
    return nullptr;
}

void wgemm_ws_launch(int c, long long n_works, const DWWork* works, const DWGroup* groups, const DWSrc* srcs, const DWDst* dsts, const double* coefs,
                     BufTable const& bufs, cudaStream_t st)
{
    if (n_works <= 0) return;
    const dim3 g((unsigned)std::min<long long>(n_works, (long long)g_occ[c] * g_sms));
    switch (c) {
#define X(cc, a, b, wm, wn, s) \
    case cc: { using Cfg = WCfg<a, b, wm, wn, s>; k_wgemm_ws<a, b, wm, wn, s><<<g, Cfg::NT, Cfg::SMEM, st>>>(works, (int)n_works, groups, srcs, dsts, coefs, bufs); break; }
        QCM_WG_VARIANTS(X)
#undef X
    }
}
