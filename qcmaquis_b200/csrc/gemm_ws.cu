// gemm_ws.cu -- grouped, variable-size FP64 GEMM for sm_100a: persistent, warp-specialised, DMMA.
//
// Replaces the block dgemm calls of the reference's contraction engine (block_matrix_algorithms.h:48-162 gemm /
// gemm_trim_left / gemm_trim_right; non-abelian/gemm.hpp:48-204) for a whole (site, direction) at once.  A work item
// is one output tile with a list of K-segments: the terms of the sum over the MPO bond index (and over source panels)
// that land in the same row unit of one symmetry sector.
//
// One CTA per SM slot runs for the whole launch and takes work items round-robin.  Its warps are specialised:
//   * four producer warps walk (work item -> K-segment -> K chunk of 16) and stages the operand tiles with cp.async
//     into a ring of shared-memory stages, in the orientation in which global memory is contiguous; completion is
//     signalled per stage through an mbarrier (cp.async.mbarrier.arrive), together with a small descriptor (alpha,
//     orientation, depth, output tile).  The ring runs across segment and work-item borders, so the operands of the
//     next output tile are in flight while the current one is being finished: short K loops (step 1 of the sigma
//     contraction has one segment per output) do not pay a pipeline fill each.
//   * the consumer warps wait on the stage's mbarrier, issue mma.sync.m8n8k4.f64 (DMMA; tcgen05 has no FP64 kind)
//     from shared memory, release the stage through a second mbarrier and write the tile (store / add / FP64 atomics
//     for split-K) when the descriptor says the work item is complete.
// Symmetry blocks are ragged: the 8x8 fragments inside the output are spread evenly over the warp grid per work item,
// fragments outside are never issued, K is consumed in steps of 4 up to the segment's real depth.
#include "qcm_dev.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int KC = 16;          // K chunk per stage
constexpr int LDK = KC + 4;     // row length of a K-major staged tile ((KC + 4) % 16 == 4: conflict-free fragment reads)
constexpr int SPAD = 4;         // padding of an M-major staged tile

struct __align__(16) StageMeta
{
    double alpha;
    long long c_off;
    int flags;          // bit 0: A staged K-major, bit 1: B staged K-major, bit 2: first chunk of a work item, bit 3: last; bits 8..: depth
    int work;           // >= 0: work item index, -1: end of stream
    int c_buf_mode;     // c_buf | mode << 8
    int ldc;
    int m0, n0, m, n;
};

// Stage a TR x kld operand tile.  g points at (row 0, k 0) of the tile.
//   row-contiguous global memory -> smem [kk][rr] (row length TR + SPAD), every cp.async of the warp covers a 256 B run
//   k-contiguous global memory   -> smem [rr][kk] (row length LDK), two 128 B runs per cp.async
// Rows in [rrem, r_ld) and depths in [krem, kld) are zero-filled (the consumers read r_ld rows and kld depths).
// The NP producer warps share a tile: warp pw takes every NP-th k row (row-contiguous case) / row pair (k-contiguous case).
template <int TR, int NP>
__device__ __forceinline__ void stage_operand(unsigned s /* shared-memory address of the stage */, const double* __restrict__ g, int ld, bool kmajor,
                                              int rrem, int r_ld, int krem, int kld, int lane, int pw)
{
    constexpr int LDS_R = TR + SPAD;
    if (!kmajor) {
        int nb[(TR + 31) / 32];
#pragma unroll
        for (int j = 0; j < (TR + 31) / 32; ++j) nb[j] = lane + 32 * j < rrem ? 8 : 0;
        const double* gp = g + (long long)pw * ld + lane;
        unsigned sp = s + (pw * LDS_R + lane) * 8;
        const long long gstep = (long long)NP * ld;
#pragma unroll 4
        for (int kk = pw; kk < kld; kk += NP) {
            const bool kv = kk < krem;
#pragma unroll
            for (int j = 0; j < (TR + 31) / 32; ++j)
                if (lane + 32 * j < r_ld) cp_async8_s(sp + 32 * j * 8, gp + 32 * j, kv ? nb[j] : 0);
            gp += gstep; sp += NP * LDS_R * 8;
        }
    } else {
        const int kk = lane & 15, r0 = (lane >> 4) + 2 * pw;
        if (kk < kld) {
            const int kb = kk < krem ? 8 : 0;
            const double* gp = g + kk + (long long)r0 * ld;
            unsigned sp = s + (r0 * LDK + kk) * 8;
            const long long gstep = 2ll * NP * ld;
#pragma unroll 4
            for (int rr = r0; rr < r_ld; rr += 2 * NP) {
                cp_async8_s(sp, gp, rr < rrem ? kb : 0);
                gp += gstep; sp += 2 * NP * LDK * 8;
            }
        }
    }
}

template <int WARPS_M, int WARPS_N, int WMT, int WNT, int STAGES>
struct GemmWsCfg
{
    static constexpr int NP = 4;     // producer warps: one warp alone cannot keep enough cp.async requests in flight
    static constexpr int TM = WARPS_M * WMT * 8, TN = WARPS_N * WNT * 8, NW = WARPS_M * WARPS_N, NT = (NW + NP) * 32;
    static constexpr int A_STAGE = (KC * (TM + SPAD) > TM * LDK) ? KC * (TM + SPAD) : TM * LDK;
    static constexpr int B_STAGE = (KC * (TN + SPAD) > TN * LDK) ? KC * (TN + SPAD) : TN * LDK;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + STAGES * (sizeof(StageMeta) + 16);
};

// CREGS > 0: the consumer warps raise their register allowance to CREGS (setmaxnreg), the producer warps drop to 40
template <int WARPS_M, int WARPS_N, int WMT, int WNT, int STAGES, int MINB, int CREGS>
__global__ void __launch_bounds__((WARPS_M * WARPS_N + 4) * 32, MINB)
k_gemm_ws(const DWork* __restrict__ works, int n_works, const DSeg* __restrict__ segs, const __grid_constant__ BufTable bufs)
{
    using Cfg = GemmWsCfg<WARPS_M, WARPS_N, WMT, WNT, STAGES>;
    constexpr int TM = Cfg::TM, TN = Cfg::TN, NW = Cfg::NW, NP = Cfg::NP, A_STAGE = Cfg::A_STAGE, B_STAGE = Cfg::B_STAGE;
    constexpr int LDA_S = TM + SPAD, LDB_S = TN + SPAD;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                               // [STAGES][A_STAGE]
    double* Bs = smem + STAGES * A_STAGE;            // [STAGES][B_STAGE]
    StageMeta* metas = reinterpret_cast<StageMeta*>(smem + STAGES * (A_STAGE + B_STAGE));
    unsigned long long* full = reinterpret_cast<unsigned long long*>(metas + STAGES);
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NP * 32 + 1); mbar_init(&empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp >= NW) {
        // ------------------------------------------------------------------ producer warps
        if (CREGS > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        const int pw = warp - NW;
        int stage = 0; unsigned phase = 0;
        auto publish = [&](StageMeta const& m) {
            if (tid == NW * 32) metas[stage] = m;
            mbar_cp_async_arrive(&full[stage]);
            if (tid == NW * 32) mbar_arrive(&full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        for (int wi = blockIdx.x; wi < n_works; wi += gridDim.x) {
            const DWork w = works[wi];
            const int tm_eff = min(TM, w.m - w.m0), tn_eff = min(TN, w.n - w.n0);
            const int tm_ld = (tm_eff + WMT * 8 - 1) / (WMT * 8) * (WMT * 8), tn_ld = (tn_eff + WNT * 8 - 1) / (WNT * 8) * (WNT * 8);   // whole warp tiles
            StageMeta m;
            m.c_off = w.c_off; m.work = wi; m.c_buf_mode = w.c_buf | (w.mode << 8); m.ldc = w.ldc;
            m.m0 = w.m0; m.n0 = w.n0; m.m = w.m; m.n = w.n;
            // last segment that contributes to this tile (segments may be smaller than the output block)
            int last_seg = -1;
            for (int s = w.seg_end - 1; s >= w.seg_begin; --s) {
                const DSeg q = segs[s];
                if (q.m - w.m0 > 0 && q.n - w.n0 > 0 && q.k > 0) { last_seg = s; break; }
            }
            int first = 4;
            if (last_seg < 0) {      // nothing to accumulate: the tile is still written (store mode) / left alone
                mbar_wait(&empty[stage], phase ^ 1);
                m.alpha = 0.; m.flags = 4 | 8;
                publish(m);
                continue;
            }
            DSeg sg = segs[w.seg_begin];
            for (int ps = w.seg_begin; ps <= last_seg; ++ps) {
                const DSeg cur = sg;
                if (ps < last_seg) sg = segs[ps + 1];        // descriptor of the next segment is in flight while this one is staged
                const int pmrem = cur.m - w.m0, pnrem = cur.n - w.n0;
                if (pmrem <= 0 || pnrem <= 0 || cur.k <= 0) continue;
                const double* __restrict__ pA = bufs.p[cur.a_buf] + cur.a_off + (cur.ta ? (long long)w.m0 * cur.lda : (long long)w.m0);
                const double* __restrict__ pB = bufs.p[cur.b_buf] + cur.b_off + (cur.tb ? (long long)w.n0 : (long long)w.n0 * cur.ldb);
                const bool a_km = cur.ta != 0, b_km = cur.tb == 0;
                m.alpha = cur.alpha;
                for (int pk0 = 0; pk0 < cur.k; pk0 += KC) {
                    const int krem = cur.k - pk0;
                    const int kld = min(KC, (krem + 3) & ~3);
                    mbar_wait(&empty[stage], phase ^ 1);
                    stage_operand<TM, NP>(smem_u32(As + stage * A_STAGE), a_km ? pA + pk0 : pA + (long long)pk0 * cur.lda, cur.lda, a_km, pmrem, tm_ld, krem, kld, lane, pw);
                    stage_operand<TN, NP>(smem_u32(Bs + stage * B_STAGE), b_km ? pB + pk0 : pB + (long long)pk0 * cur.ldb, cur.ldb, b_km, pnrem, tn_ld, krem, kld, lane, pw);
                    const bool last = ps == last_seg && pk0 + KC >= cur.k;
                    m.flags = (a_km ? 1 : 0) | (b_km ? 2 : 0) | first | (last ? 8 : 0) | (kld << 8);
                    first = 0;
                    publish(m);
                }
            }
        }
        // end of stream
        mbar_wait(&empty[stage], phase ^ 1);
        StageMeta m;
        m.alpha = 0.; m.c_off = 0; m.flags = 0; m.work = -1; m.c_buf_mode = 0; m.ldc = 0; m.m0 = m.n0 = m.m = m.n = 0;
        publish(m);
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    if (CREGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(CREGS > 0 ? CREGS : 24));
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int fr = lane >> 2, fk = lane & 3;
    const int row0 = wm * WMT * 8, col0 = wn * WNT * 8;      // this warp's WMT x WNT fragments inside the tile
    double acc[WMT][WNT][2];
    bool active = false;                                      // warp tile intersects the output
    int stage = 0; unsigned phase = 0;
    for (;;) {
        mbar_wait(&full[stage], phase);
        const StageMeta sm = metas[stage];
        if (sm.work < 0) break;
        const int fl = sm.flags;
        if (fl & 4) {
#pragma unroll
            for (int i = 0; i < WMT; ++i)
#pragma unroll
                for (int j = 0; j < WNT; ++j) acc[i][j][0] = acc[i][j][1] = 0.;
            active = sm.m0 + row0 < sm.m && sm.n0 + col0 < sm.n;
        }
        const int k4n = fl >> 10;
        if (k4n > 0 && active) {
            const double alpha = sm.alpha;
            const int a_base = (fl & 1) ? (row0 + fr) * LDK + fk : fk * LDA_S + row0 + fr;
            const int a_ti = (fl & 1) ? 8 * LDK : 8, a_tk = (fl & 1) ? 4 : 4 * LDA_S;
            const int b_base = (fl & 2) ? (col0 + fr) * LDK + fk : fk * LDB_S + col0 + fr;
            const int b_tj = (fl & 2) ? 8 * LDK : 8, b_tk = (fl & 2) ? 4 : 4 * LDB_S;
            const double* as = As + stage * A_STAGE + a_base;
            const double* bs = Bs + stage * B_STAGE + b_base;
            if (k4n == KC / 4) {
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    double a[WMT], b[WNT];
#pragma unroll
                    for (int i = 0; i < WMT; ++i) a[i] = alpha * as[i * a_ti + k4 * a_tk];
#pragma unroll
                    for (int j = 0; j < WNT; ++j) b[j] = bs[j * b_tj + k4 * b_tk];
#pragma unroll
                    for (int i = 0; i < WMT; ++i)
#pragma unroll
                        for (int j = 0; j < WNT; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            } else {
                for (int k4 = 0; k4 < k4n; ++k4) {
                    double a[WMT], b[WNT];
#pragma unroll
                    for (int i = 0; i < WMT; ++i) a[i] = alpha * as[i * a_ti + k4 * a_tk];
#pragma unroll
                    for (int j = 0; j < WNT; ++j) b[j] = bs[j * b_tj + k4 * b_tk];
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
            // ---- epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1})
            const int mode = sm.c_buf_mode >> 8;
            const int r0 = sm.m0 + row0 + fr, c0 = sm.n0 + col0 + 2 * fk;
            double* __restrict__ q0 = bufs.p[sm.c_buf_mode & 0xff] + sm.c_off + r0 + (long long)c0 * sm.ldc;
            const long long cstep = 8ll * sm.ldc;
            if (sm.m0 + row0 + WMT * 8 <= sm.m && sm.n0 + col0 + WNT * 8 <= sm.n) {     // warp tile fully inside
#pragma unroll
                for (int j = 0; j < WNT; ++j) {
                    double* qj = q0 + j * cstep;
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int i = 0; i < WMT; ++i) {
                            double* q = qj + e * sm.ldc + i * 8;
                            if (mode == 0) *q = acc[i][j][e];
                            else if (mode == 1) *q += acc[i][j][e];
                            else atomicAdd(q, acc[i][j][e]);
                        }
                }
            } else {
#pragma unroll
                for (int j = 0; j < WNT; ++j) {
                    double* qj = q0 + j * cstep;
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int i = 0; i < WMT; ++i) {
                            if (r0 + i * 8 < sm.m && c0 + j * 8 + e < sm.n) {
                                double* q = qj + e * sm.ldc + i * 8;
                                if (mode == 0) *q = acc[i][j][e];
                                else if (mode == 1) *q += acc[i][j][e];
                                else atomicAdd(q, acc[i][j][e]);
                            }
                        }
                }
            }
        }
    }
}

// ---- variant table ----------------------------------------------------------------------------------------------
//   X(index, WARPS_M, WARPS_N, WMT, WNT, STAGES, MINB, CREGS)      tile = (WARPS_M * WMT * 8) x (WARPS_N * WNT * 8)
#define QCM_WS_VARIANTS(X) \
    X(0, 4, 2, 4, 8, 5, 1, 232) /* 128 x 128 */ \
    X(1, 2, 4, 4, 4, 6, 1, 0)   /*  64 x 128 */ \
    X(2, 4, 2, 4, 4, 6, 1, 0)   /* 128 x  64 */ \
    X(3, 1, 8, 4, 2, 7, 1, 0)   /*  32 x 128 */ \
    X(4, 8, 1, 2, 4, 7, 1, 0)   /* 128 x  32 */ \
    X(5, 1, 8, 2, 2, 8, 1, 0)   /*  16 x 128 */ \
    X(6, 8, 1, 2, 2, 8, 1, 0)   /* 128 x  16 */ \
    X(7, 1, 8, 1, 2, 8, 1, 0)   /*   8 x 128 */ \
    X(8, 8, 1, 2, 1, 8, 1, 0)   /* 128 x   8 */ \
    X(9, 2, 4, 4, 2, 8, 1, 0)   /*  64 x  64 */ \
    X(10, 2, 2, 2, 2, 6, 2, 0)  /*  32 x  32 */ \
    X(11, 2, 2, 1, 1, 6, 2, 0)  /*  16 x  16 */ \
    X(12, 4, 2, 2, 2, 8, 1, 0)  /*  64 x  32 */ \
    X(13, 2, 4, 2, 2, 8, 1, 0)  /*  32 x  64 */ \
    X(14, 8, 1, 1, 2, 8, 1, 0)  /*  64 x  16 */ \
    X(15, 1, 8, 2, 1, 8, 1, 0)  /*  16 x  64 */

constexpr int kNumWs = 16;
GemmWsVariant g_var[kNumWs];
int g_occ[kNumWs];
int g_sms = 0;

}  // namespace

int gemm_ws_num_variants() { return kNumWs; }
GemmWsVariant gemm_ws_variant(int v) { return g_var[v]; }

const char* gemm_ws_init(int sm_count)
{
    g_sms = sm_count;
    cudaError_t e;
#define X(v, a, b, c, d, s, mb, cr) \
    { using Cfg = GemmWsCfg<a, b, c, d, s>; \
      e = cudaFuncSetAttribute(k_gemm_ws<a, b, c, d, s, mb, cr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM); \
      if (e != cudaSuccess) return cudaGetErrorString(e); \
      int occ = 0; \
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gemm_ws<a, b, c, d, s, mb, cr>, Cfg::NT, Cfg::SMEM); \
      if (e != cudaSuccess) return cudaGetErrorString(e); \
      if (occ < 1) return "k_gemm_ws variant does not fit on an SM"; \
      g_occ[v] = occ; g_var[v] = GemmWsVariant{Cfg::TM, Cfg::TN, Cfg::NT, 1.0}; \
      if (getenv("QCM_DEBUG")) fprintf(stderr, "gemm_ws variant %d: tile %d x %d, %d threads, %zu B smem, %d CTAs/SM\n", v, Cfg::TM, Cfg::TN, Cfg::NT, (size_t)Cfg::SMEM, occ); }
    QCM_WS_VARIANTS(X)
#undef X
    return nullptr;
}

int gemm_ws_grid(int v, long long n_works) { return (int)std::min<long long>(n_works, (long long)g_occ[v] * g_sms); }

void gemm_ws_launch(int v, long long n_works, const DWork* works, const DSeg* segs, BufTable const& bufs, cudaStream_t st)
{
    if (n_works <= 0) return;
    const dim3 g((unsigned)gemm_ws_grid(v, n_works));
    switch (v) {
#define X(vv, a, b, c, d, s, mb, cr) \
    case vv: { using Cfg = GemmWsCfg<a, b, c, d, s>; k_gemm_ws<a, b, c, d, s, mb, cr><<<g, Cfg::NT, Cfg::SMEM, st>>>(works, (int)n_works, segs, bufs); break; }
        QCM_WS_VARIANTS(X)
#undef X
    }
}
