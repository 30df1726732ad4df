// plan::Plan -> qcm_plan_desc (the task-array form of include/qcm_b200.h).  Used by qcm::GpuEngine (host objects above the C
// ABI) and by the descriptor entry points inside the library (csrc/plan_capi.cpp).
#pragma once
#include "../../../include/qcm_b200.h"
#include "plan.hpp"
#include <cstring>

namespace qcm {

struct PlanDescHolder
{
    struct WaveStore { std::vector<qcm_gemm_out> to, co; std::vector<qcm_gemm_seg> ts, cs; std::vector<qcm_w_group> wg; std::vector<qcm_w_dst> wd; std::vector<qcm_w_src> wsrc; };
    std::vector<WaveStore> store;
    std::vector<qcm_wave_desc> waves;
    std::vector<qcm_gemm_out> po; std::vector<qcm_gemm_seg> ps;
    std::vector<qcm_copy_task> copies;
    qcm_plan_desc d;

    static void cvt(plan::GemmList const& g, std::vector<qcm_gemm_out>& outs, std::vector<qcm_gemm_seg>& segs)
    {
        outs.resize(g.outs.size()); segs.resize(g.segs.size());
        // 10^6 records per list at cfg3: converted on all cores (the caller sits between two device phases of a sweep)
#pragma omp parallel for schedule(static) if (g.outs.size() > 65536)
        for (size_t i = 0; i < g.outs.size(); ++i) {
            plan::Out const& o = g.outs[i];
            outs[i] = qcm_gemm_out{qcm_ref{o.C.buf, 0, o.C.off}, o.ldc, o.m, o.n, o.seg_begin, o.seg_end, 0};
        }
#pragma omp parallel for schedule(static) if (g.segs.size() > 65536)
        for (size_t i = 0; i < g.segs.size(); ++i) {
            plan::Seg const& s = g.segs[i];
            segs[i] = qcm_gemm_seg{qcm_ref{s.A.buf, 0, s.A.off}, qcm_ref{s.B.buf, 0, s.B.off}, s.lda, s.ldb, s.m, s.n, s.k, s.ta, s.tb, 0, s.alpha};
        }
    }
    // P must outlive the holder's use (the W coefficient tables are referenced, not copied)
    void fill(plan::Plan const& P, int64_t left_elems, int64_t right_elems)
    {
        store.assign(P.waves.size(), WaveStore()); waves.resize(P.waves.size());
        for (size_t w = 0; w < P.waves.size(); ++w) {
            plan::Wave const& W = P.waves[w];
            WaveStore& S = store[w];
            cvt(W.t_gemm, S.to, S.ts);
            cvt(W.close_gemm, S.co, S.cs);
            plan::WList const& wl = W.w_groups;
            S.wg.resize(wl.groups.size()); S.wd.resize(wl.dsts.size()); S.wsrc.resize(wl.srcs.size());
            for (size_t i = 0; i < S.wg.size(); ++i) {
                plan::WGroup const& g = wl.groups[i];
                S.wg[i] = qcm_w_group{g.rows, g.cols, g.n_src, g.n_dst, g.ng, g.src_begin, g.dst_begin, g.cls, g.coef_begin};
            }
#pragma omp parallel for schedule(static) if (S.wd.size() > 65536)
            for (size_t i = 0; i < S.wd.size(); ++i) S.wd[i] = qcm_w_dst{qcm_ref{wl.dsts[i].dst.buf, 0, wl.dsts[i].dst.off}, wl.dsts[i].ldd, 0};
#pragma omp parallel for schedule(static) if (S.wsrc.size() > 65536)
            for (size_t i = 0; i < S.wsrc.size(); ++i) S.wsrc[i] = qcm_w_src{qcm_ref{wl.srcs[i].src.buf, 0, wl.srcs[i].src.off}, wl.srcs[i].lds, 0};
            waves[w] = qcm_wave_desc{S.to.data(), (int64_t)S.to.size(), S.ts.data(), (int64_t)S.ts.size(),
                                     S.wg.data(), (int64_t)S.wg.size(), S.wsrc.data(), (int64_t)S.wsrc.size(), S.wd.data(), (int64_t)S.wd.size(),
                                     wl.coefs.data(), (int64_t)wl.coefs.size(),
                                     S.co.data(), (int64_t)S.co.size(), S.cs.data(), (int64_t)S.cs.size(), W.y_elems, W.t_elems,
                                     W.x_chunk, W.x_zero ? 1 : 0, 0};
        }
        cvt(P.persistent_t, po, ps);
        copies.resize(P.pre_copies.size());
        for (size_t i = 0; i < copies.size(); ++i) {
            plan::CopyTask const& c = P.pre_copies[i];
            copies[i] = qcm_copy_task{qcm_ref{c.src.buf, 0, c.src.off}, qcm_ref{c.dst.buf, 0, c.dst.off}, c.rows, c.cols, c.lds, c.ldd};
        }
        std::memset(&d, 0, sizeof(d));
        d.kind = P.kind; d.n_waves = (int32_t)waves.size();
        d.pre_copies = copies.data(); d.n_pre_copies = (int64_t)copies.size();
        d.p_outs = po.data(); d.n_p_outs = (int64_t)po.size(); d.p_segs = ps.data(); d.n_p_segs = (int64_t)ps.size();
        d.waves = waves.data();
        d.elems[QCM_BUF_KET_LP] = P.ket_lp_elems; d.elems[QCM_BUF_KET_RP] = P.ket_rp_elems;
        d.elems[QCM_BUF_LEFT] = left_elems; d.elems[QCM_BUF_RIGHT] = right_elems;
        d.elems[QCM_BUF_T] = P.t_elems_max; d.elems[QCM_BUF_TP] = P.tp_elems; d.elems[QCM_BUF_Y] = P.y_elems_max;
        d.elems[QCM_BUF_OUT] = out_elems(P); d.elems[QCM_BUF_BRA_LP] = P.bra_lp_elems; d.elems[QCM_BUF_BRA_RP] = P.bra_rp_elems;
        d.flops = P.flops(); d.bytes = P.bytes_algorithmic;
        d.rank = P.rank; d.world = P.world;
    }
    static int64_t out_elems(plan::Plan const& P) { return (P.kind == 0 || P.kind == 3) ? P.out_tensor.total : P.out_boundary.total; }
};

} // namespace qcm
