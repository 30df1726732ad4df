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

inline void run_plan(plan::Plan const& P, Bufs& B)
{
    B.b[plan::BUF_KET_RP].assign((size_t)P.ket_rp_elems, 0.);
    B.b[plan::BUF_BRA_RP].assign((size_t)P.bra_rp_elems, 0.);
    B.b[plan::BUF_T].assign((size_t)P.t_elems_max, 0.);
    B.b[plan::BUF_TP].assign((size_t)P.tp_elems, 0.);
    B.b[plan::BUF_Y].assign((size_t)P.y_elems_max, 0.);
    for (auto const& c : P.pre_copies) {
        const double* s = B.p(c.src); double* d = B.p(c.dst);
        for (int j = 0; j < c.cols; ++j) for (int i = 0; i < c.rows; ++i) d[i + (size_t)j * c.ldd] = s[i + (size_t)j * c.lds];
    }
    run_gemm(P.persistent_t, B, false);
    for (auto const& W : P.waves) {
        // poison T so that a read of a product that was not computed in this wave is caught
        std::fill(B.b[plan::BUF_T].begin(), B.b[plan::BUF_T].end(), std::nan(""));
        run_gemm(W.t_gemm, B, false);
        // Y is never zero-filled: every element the closing products read was written by the W pass of this wave
        // (the diagonal_hamiltonian plan, kind 3, is the exception: rows of V without contributions are zero, qcm_hdiag)
        std::fill(B.b[plan::BUF_Y].begin(), B.b[plan::BUF_Y].end(), P.kind == 3 ? 0. : std::nan(""));
        for (auto const& G : W.w_groups.groups)
            for (int d = 0; d < G.n_dst; ++d) {
                plan::WDst const& D = W.w_groups.dsts[G.dst_begin + d];
                double* dst = B.p(D.dst);
                for (int j = 0; j < G.cols; ++j)
                    for (int i = 0; i < G.rows; ++i) {
                        double acc = 0.;
                        for (int u = 0; u < G.n_src; ++u) {
                            plan::WSrc const& q = W.w_groups.srcs[G.src_begin + u];
                            acc += W.w_groups.coefs[G.coef_begin + (size_t)u * G.ng + d] * B.p(q.src)[i + (size_t)j * q.lds];
                        }
                        dst[i + (size_t)j * D.ldd] = acc;
                    }
            }
        run_gemm(W.close_gemm, B, true);
    }
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
        std::vector<double> sum; plan::Layout out;
        for (int r = 0; r < world; ++r) {
            plan::Planner pl(symm, mpo, isHermitian, r, world, budget);
            plan::Plan P = pl.plan_sigma(desc_of(ket), ll, rl);
            Bufs B;
            B.b[plan::BUF_KET_LP] = flat(ket.data()); B.b[plan::BUF_LEFT] = flat(left); B.b[plan::BUF_RIGHT] = flat(right);
            B.b[plan::BUF_OUT].assign((size_t)P.out_tensor.total, 0.);
            run_plan(P, B);
            if (r == 0) { sum = B.b[plan::BUF_OUT]; out = P.out_tensor; }
            else {
                if (!(out.basis == P.out_tensor.basis)) throw std::runtime_error("rank plans disagree on the sigma structure");
                for (size_t i = 0; i < sum.size(); ++i) sum[i] += B.b[plan::BUF_OUT][i];
            }
            last_flops = P.flops(); last_waves = P.waves.size();
        }
        return MPSTensor(ket.site_dim(), ket.row_dim(), ket.col_dim(), unflat(out, sum, 0), LeftPaired, true);
    }
    Boundary step(int kind, MPSTensor const& bra, MPSTensor const& ket, Boundary const& in, MPOTensor const& mpo, bool isHermitian)
    {
        bra.make_left_paired(); ket.make_left_paired();
        plan::BoundaryLayout il = layout_of(in);
        std::vector<double> sum; plan::BoundaryLayout out;
        for (int r = 0; r < world; ++r) {
            plan::Planner pl(symm, mpo, isHermitian, r, world, budget);
            plan::Plan P = kind == 1 ? pl.plan_left_step(desc_of(bra), desc_of(ket), il) : pl.plan_right_step(desc_of(bra), desc_of(ket), il);
            Bufs B;
            B.b[plan::BUF_KET_LP] = flat(ket.data()); B.b[plan::BUF_BRA_LP] = flat(bra.data());
            B.b[kind == 1 ? plan::BUF_LEFT : plan::BUF_RIGHT] = flat(in);
            B.b[plan::BUF_OUT].assign((size_t)P.out_boundary.total, 0.);
            run_plan(P, B);
            if (r == 0) { sum = B.b[plan::BUF_OUT]; out = P.out_boundary; }
            else for (size_t i = 0; i < sum.size(); ++i) sum[i] += B.b[plan::BUF_OUT][i];
        }
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
    Boundary overlap_mpo_left_step(MPSTensor const& bra, MPSTensor const& ket, Boundary const& left, MPOTensor const& mpo, bool h = true) override { return step(1, bra, ket, left, mpo, h); }
    Boundary overlap_mpo_right_step(MPSTensor const& bra, MPSTensor const& ket, Boundary const& right, MPOTensor const& mpo, bool h = true) override { return step(2, bra, ket, right, mpo, h); }
    double last_flops = 0; size_t last_waves = 0;

private:
    SymmKind symm; int world; int64_t budget;
};

} // namespace qcmtest
