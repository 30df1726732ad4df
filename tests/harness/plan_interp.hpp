// TEST INFRASTRUCTURE -- a plain-loop CPU interpreter of plan::Plan.
// It exists so that the host-side schedule builder (qcm/plan.hpp) can be checked against the oracle on a
// machine without a GPU, including the multi-rank sharding (each rank's plan is executed and the partial
// results are summed, which is what the NCCL allreduce does on the device).  It is never linked into the
// product libraries; the product path executes plans on the GPU only.
#pragma once
#include "qcm/engine_iface.hpp"
#include "qcm/plan.hpp"

namespace qcmtest {
using namespace qcm;

struct Bufs { std::vector<double> b[plan::BUF_COUNT]; double* p(plan::Ref r) { return b[r.buf].data() + r.off; } };

inline void run_gemm(plan::GemmList const& g, Bufs& B, bool accumulate)
{
    for (auto const& o : g.outs) {
        double* C = B.p(o.C);
        if (!accumulate)
            for (int j = 0; j < o.n; ++j) for (int i = 0; i < o.m; ++i) C[i + (size_t)j * o.ldc] = 0.;
        for (int s = o.seg_begin; s < o.seg_end; ++s) {
            plan::Seg const& sg = g.segs[s];
            const double* A = B.p(sg.A); const double* Bm = B.p(sg.B);
            for (int j = 0; j < sg.n; ++j)
                for (int i = 0; i < sg.m; ++i) {
                    double acc = 0.;
                    for (int k = 0; k < sg.k; ++k) {
                        double a = sg.ta ? A[k + (size_t)i * sg.lda] : A[i + (size_t)k * sg.lda];
                        double b = sg.tb ? Bm[j + (size_t)k * sg.ldb] : Bm[k + (size_t)j * sg.ldb];
                        acc += a * b;
                    }
                    C[i + (size_t)j * o.ldc] += sg.alpha * acc;
                }
        }
    }
}

inline void run_w(plan::WList const& wl, Bufs& B)
{
    for (auto const& G : wl.groups)
        for (int d = 0; d < G.n_dst; ++d) {
            plan::WDst const& D = wl.dsts[G.dst_begin + d];
            double* dst = B.p(D.dst);
            for (int j = 0; j < G.cols; ++j)
                for (int i = 0; i < G.rows; ++i) {
                    double acc = 0.;
                    for (int u = 0; u < G.n_src; ++u) {
                        plan::WSrc const& q = wl.srcs[G.src_begin + u];
                        acc += wl.coefs[G.coef_begin + (size_t)u * G.ng + d] * B.p(q.src)[i + (size_t)j * q.lds];
                    }
                    dst[i + (size_t)j * D.ldd] = acc;
                }
        }
}
// workspaces, pairing reshapes, resident step-1 products
inline void run_pre(plan::Plan const& P, Bufs& B)
{
    B.b[plan::BUF_KET_RP].assign((size_t)P.ket_rp_elems, 0.);
    B.b[plan::BUF_BRA_RP].assign((size_t)P.bra_rp_elems, 0.);
    B.b[plan::BUF_T].assign((size_t)P.t_elems_max, 0.);
    B.b[plan::BUF_TP].assign((size_t)P.tp_elems, 0.);
    B.b[plan::BUF_Y].assign((size_t)P.y_elems_max, std::nan(""));
    for (auto const& c : P.pre_copies) {
        const double* s = B.p(c.src); double* d = B.p(c.dst);
        for (int j = 0; j < c.cols; ++j) for (int i = 0; i < c.rows; ++i) d[i + (size_t)j * c.ldd] = s[i + (size_t)j * c.lds];
    }
    run_gemm(P.persistent_t, B, false);
}
inline int64_t exchange_elems(plan::Plan const& P) { return (!P.waves.empty() && P.waves[0].x_chunk > 0) ? P.waves[0].x_chunk * P.world : 0; }
// the W pass of the exchange wave: this rank's partial sums of the exchanged panels
inline void run_exchange_w(plan::Plan const& P, Bufs& B)
{
    plan::Wave const& W = P.waves[0];
    // without x_zero every element of the region must be written by the W pass: NaN poison catches a slot that is not
    std::fill(B.b[plan::BUF_Y].begin(), B.b[plan::BUF_Y].begin() + exchange_elems(P), W.x_zero ? 0. : std::nan(""));
    run_w(W.w_groups, B);
}
// the waves a rank runs on its own (everything but the exchange wave), then the closing products of its exchange chunk
inline void run_local(plan::Plan const& P, Bufs& B)
{
    const int64_t y0 = exchange_elems(P);
    for (size_t w = y0 > 0 ? 1 : 0; w < P.waves.size(); ++w) {
        plan::Wave const& W = P.waves[w];
        if (W.x_chunk > 0) throw std::runtime_error("exchange wave is not waves[0]");
        // poison T so that a read of a product that was not computed in this wave is caught
        std::fill(B.b[plan::BUF_T].begin(), B.b[plan::BUF_T].end(), std::nan(""));
        run_gemm(W.t_gemm, B, false);
        // Y is never zero-filled: every element the closing products read was written by the W pass of this wave
        // (the diagonal_hamiltonian plan, kind 3, is the exception: rows of V without contributions are zero, qcm_hdiag)
        std::fill(B.b[plan::BUF_Y].begin() + y0, B.b[plan::BUF_Y].end(), P.kind == 3 ? 0. : std::nan(""));
        run_w(W.w_groups, B);
        run_gemm(W.close_gemm, B, true);
    }
    if (y0 > 0) run_gemm(P.waves[0].close_gemm, B, true);
}
inline void run_plan(plan::Plan const& P, Bufs& B)
{
    if (exchange_elems(P) > 0) throw std::runtime_error("run_plan: a plan with an exchange wave needs all ranks (run_plans)");
    run_pre(P, B);
    run_local(P, B);
}
// all ranks of a sharded contraction in lockstep; the exchange region is reduce-scattered the way ncclReduceScatter does it:
// rank r receives the sum of chunk r, the other chunks of its region hold garbage afterwards (poisoned here)
inline void run_plans(std::vector<plan::Plan> const& Ps, std::vector<Bufs>& Bs)
{
    const size_t N = Ps.size();
    for (size_t r = 0; r < N; ++r) run_pre(Ps[r], Bs[r]);
    const int64_t xe = exchange_elems(Ps[0]);
    for (size_t r = 0; r < N; ++r) if (exchange_elems(Ps[r]) != xe) throw std::runtime_error("ranks disagree on the exchange region");
    if (xe > 0) {
        const int64_t C = Ps[0].waves[0].x_chunk;
        for (size_t r = 0; r < N; ++r) run_exchange_w(Ps[r], Bs[r]);
        std::vector<double> sum((size_t)xe, 0.);
        for (size_t r = 0; r < N; ++r) for (int64_t i = 0; i < xe; ++i) sum[(size_t)i] += Bs[r].b[plan::BUF_Y][(size_t)i];
        for (size_t r = 0; r < N; ++r)
            for (int64_t i = 0; i < xe; ++i) Bs[r].b[plan::BUF_Y][(size_t)i] = (i / C == (int64_t)r) ? sum[(size_t)i] : std::nan("");
    }
    for (size_t r = 0; r < N; ++r) run_local(Ps[r], Bs[r]);
}

class InterpEngine : public EngineIface
{
public:
    explicit InterpEngine(SymmKind s, int world_ = 1, int64_t budget_ = (int64_t)1 << 28) : symm(s), world(world_), budget(budget_) {}

    static plan::TensorDesc desc_of(MPSTensor const& t) { t.make_left_paired(); return plan::TensorDesc{t.site_dim(), t.row_dim(), t.col_dim(), t.data().basis()}; }
    static std::vector<double> flat(block_matrix const& m) { std::vector<double> f; for (size_t k = 0; k < m.n_blocks(); ++k) f.insert(f.end(), m[k].v.begin(), m[k].v.end()); return f; }
    static plan::BoundaryLayout layout_of(Boundary const& b)
    {
        std::vector<DualIndex> bases(b.aux_dim());
        for (size_t k = 0; k < b.aux_dim(); ++k) bases[k] = b[k].basis();
        plan::BoundaryLayout L; L.assign(bases); return L;
    }
    static std::vector<double> flat(Boundary const& b) { std::vector<double> f; for (size_t k = 0; k < b.aux_dim(); ++k) { auto g = flat(b[k]); f.insert(f.end(), g.begin(), g.end()); } return f; }
    static block_matrix unflat(plan::Layout const& L, std::vector<double> const& f, int64_t base)
    {
        block_matrix r;
        for (size_t k = 0; k < L.basis.size(); ++k) {
            Matrix m(L.basis[k].ls, L.basis[k].rs);
            std::copy(f.begin() + L.off[k] - base, f.begin() + L.off[k] - base + m.v.size(), m.v.begin());
            r.insert_block(std::move(m), L.basis[k].lc, L.basis[k].rc);
        }
        return r;
    }

    MPSTensor site_hamil2(MPSTensor ket, Boundary const& left, Boundary const& right, MPOTensor const& mpo, bool isHermitian = true) override
    {
        ket.make_left_paired();
        plan::BoundaryLayout ll = layout_of(left), rl = layout_of(right);
        std::vector<plan::Plan> Ps; std::vector<Bufs> Bs((size_t)world);
        for (int r = 0; r < world; ++r) {
            plan::Planner pl(symm, mpo, isHermitian, r, world, budget);
            Ps.push_back(pl.plan_sigma(desc_of(ket), ll, rl));
            Bufs& B = Bs[(size_t)r];
            B.b[plan::BUF_KET_LP] = flat(ket.data()); B.b[plan::BUF_LEFT] = flat(left); B.b[plan::BUF_RIGHT] = flat(right);
            B.b[plan::BUF_OUT].assign((size_t)Ps.back().out_tensor.total, 0.);
        }
        run_plans(Ps, Bs);
        std::vector<double> sum = Bs[0].b[plan::BUF_OUT]; plan::Layout out = Ps[0].out_tensor;
        last_exchange_elems = exchange_elems(Ps[0]);
        last_exec_close = 0;
        for (int r = 0; r < world; ++r) {
            plan::Plan const& P = Ps[(size_t)r];
            last_exec_close += P.exec_close;
            if (r > 0) {
                if (!(out.basis == P.out_tensor.basis)) throw std::runtime_error("rank plans disagree on the sigma structure");
                for (size_t i = 0; i < sum.size(); ++i) sum[i] += Bs[(size_t)r].b[plan::BUF_OUT][i];
            }
            last_flops = P.flops(); last_waves = P.waves.size();
        }
        return MPSTensor(ket.site_dim(), ket.row_dim(), ket.col_dim(), unflat(out, sum, 0), LeftPaired, true);
    }
    Boundary step(int kind, MPSTensor const& bra, MPSTensor const& ket, Boundary const& in, MPOTensor const& mpo, bool isHermitian)
    {
        bra.make_left_paired(); ket.make_left_paired();
        plan::BoundaryLayout il = layout_of(in);
        std::vector<plan::Plan> Ps; std::vector<Bufs> Bs((size_t)world);
        for (int r = 0; r < world; ++r) {
            plan::Planner pl(symm, mpo, isHermitian, r, world, budget);
            Ps.push_back(kind == 1 ? pl.plan_left_step(desc_of(bra), desc_of(ket), il) : pl.plan_right_step(desc_of(bra), desc_of(ket), il));
            Bufs& B = Bs[(size_t)r];
            B.b[plan::BUF_KET_LP] = flat(ket.data()); B.b[plan::BUF_BRA_LP] = flat(bra.data());
            B.b[kind == 1 ? plan::BUF_LEFT : plan::BUF_RIGHT] = flat(in);
            B.b[plan::BUF_OUT].assign((size_t)Ps.back().out_boundary.total, 0.);
        }
        run_plans(Ps, Bs);
        std::vector<double> sum = Bs[0].b[plan::BUF_OUT]; plan::BoundaryLayout out = Ps[0].out_boundary;
        for (int r = 1; r < world; ++r) for (size_t i = 0; i < sum.size(); ++i) sum[i] += Bs[(size_t)r].b[plan::BUF_OUT][i];
        Boundary ret; ret.resize(out.aux_dim());
        for (size_t b = 0; b < out.aux_dim(); ++b) ret[b] = unflat(out.b[b], sum, 0);
        return ret;
    }
    block_matrix diagonal_hamiltonian(Boundary const& left, Boundary const& right, MPOTensor const& mpo, MPSTensor const& x)
    {
        plan::BoundaryLayout ll = layout_of(left), rl = layout_of(right);
        plan::Planner pl(symm, mpo, true, 0, 1, budget);
        plan::Plan P = pl.plan_hdiag(desc_of(x), ll, rl);
        Bufs B;
        B.b[plan::BUF_LEFT] = flat(left); B.b[plan::BUF_RIGHT] = flat(right);
        B.b[plan::BUF_OUT].assign((size_t)P.out_tensor.total, 0.);
        run_plan(P, B);
        return unflat(P.out_tensor, B.b[plan::BUF_OUT], 0);
    }
    block_matrix noise(bool left_side, MPSTensor const& mps, Boundary const& in, MPOTensor const& mpo)
    {
        mps.make_left_paired();
        plan::BoundaryLayout il = layout_of(in);
        plan::Planner pl(symm, mpo, true, 0, 1, budget);
        plan::Plan P = left_side ? pl.plan_noise_left(desc_of(mps), il) : pl.plan_noise_right(desc_of(mps), il);
        Bufs B;
        B.b[plan::BUF_KET_LP] = flat(mps.data());
        B.b[left_side ? plan::BUF_LEFT : plan::BUF_RIGHT] = flat(in);
        B.b[plan::BUF_OUT].assign((size_t)P.out_boundary.total, 0.);
        run_plan(P, B);
        return unflat(P.out_boundary.b[0], B.b[plan::BUF_OUT], 0);
    }
    block_matrix noise_left(MPSTensor const& mps, Boundary const& left, MPOTensor const& mpo) override { return noise(true, mps, left, mpo); }
    block_matrix noise_right(MPSTensor const& mps, Boundary const& right, MPOTensor const& mpo) override { return noise(false, mps, right, mpo); }
    Boundary overlap_mpo_left_step(MPSTensor const& bra, MPSTensor const& ket, Boundary const& left, MPOTensor const& mpo, bool h = true) override { return step(1, bra, ket, left, mpo, h); }
    Boundary overlap_mpo_right_step(MPSTensor const& bra, MPSTensor const& ket, Boundary const& right, MPOTensor const& mpo, bool h = true) override { return step(2, bra, ket, right, mpo, h); }
    double last_flops = 0, last_exec_close = 0; size_t last_waves = 0; int64_t last_exchange_elems = 0;

private:
    SymmKind symm; int world; int64_t budget;
};

} // namespace qcmtest
