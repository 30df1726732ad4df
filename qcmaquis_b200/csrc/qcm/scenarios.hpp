// Host-side drivers around the hot path, written against EngineIface so that the same code runs on the B200
// engine and on the CPU checker: boundary chains (BoundaryPropagator.h:36-150, optimize.h:119-165), the
// sigma-energy identity of dmrg/tests/test_mps_mpo_ops/test_siteproblem.cpp:38-95, dense effective
// Hamiltonians for exact small cases, and comparison helpers.
#pragma once
#include "engine_iface.hpp"
#include "models.hpp"

extern "C" void scipy_dsyev_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w,
                             double* work, const int* lwork, int* info);

namespace qcm {

struct Problem
{
    ModelParams params;
    std::shared_ptr<ModelBase> model;
    MPO mpo;
    MPS mps;
    std::vector<Boundary> left, right;   // left[p]: everything left of site p; right[p]: everything right of site p-1
    std::map<int, MPOTensor> ts_mpo;

    SymmKind symm() const { return params.symm; }
    std::vector<int> site_types() const { return model->lat.irreps; }
    Index const& phys(int p) const { return model->phys_dim(model->lat.type(p)); }

    void build_model() { model = make_model(params); }
    void build_mpo() { TaggedMPOMaker maker(*model); mpo = maker.create_mpo(); }
    void init_mps(size_t Mmax, bool fillrand, double val, unsigned seed)
    {
        mps = make_mps(symm(), site_types(), model->phys_indices, model->total_charge, Mmax, fillrand, val, seed);
    }
    MPOTensor const& twosite_mpo(int p)
    {
        auto it = ts_mpo.find(p);
        if (it == ts_mpo.end()) it = ts_mpo.emplace(p, make_twosite_mpo(symm(), mpo[p], mpo[p + 1], phys(p), phys(p + 1))).first;
        return it->second;
    }
    // all boundaries, as BoundaryPropagator's constructor does
    void build_boundaries(EngineIface& eng, int left_upto = -1, int right_downto = -1)
    {
        int L = (int)mps.size();
        if (left_upto < 0) left_upto = L;
        if (right_downto < 0) right_downto = 0;
        left.assign(L + 1, Boundary()); right.assign(L + 1, Boundary());
        left[0] = mps.left_boundary();
        for (int i = 0; i < left_upto; ++i) left[i + 1] = eng.overlap_mpo_left_step(mps[i], mps[i], left[i], mpo[i]);
        right[L] = mps.right_boundary();
        for (int i = L - 1; i >= right_downto; --i) right[i] = eng.overlap_mpo_right_step(mps[i], mps[i], right[i + 1], mpo[i]);
    }
};

// two-site tensor basis: physical index of the fused site (twositetensor.hpp:23-31; ts_reduction.h:50,147)
inline MPSTensor make_twosite_tensor(Index const& phys1, Index const& phys2, Index const& left_i, Index const& right_i, std::function<double()> gen)
{
    return MPSTensor(phys1 * phys2, left_i, right_i, gen);
}

struct DiffReport { double max_abs = 0, ref_norm = 0, diff_norm = 0; int structure_equal = 1; };

inline void accumulate_diff(block_matrix const& a, block_matrix const& ref, DiffReport& r)
{
    if (!(a.basis() == ref.basis())) {
        r.structure_equal = 0;
        // still measure what can be matched
    }
    for (size_t k = 0; k < ref.n_blocks(); ++k) {
        size_t j = a.find_block(ref.basis().left_charge(k), ref.basis().right_charge(k));
        Matrix const& m = ref[k];
        for (double x : m.v) r.ref_norm += x * x;
        if (j == a.n_blocks() || a[j].rows != m.rows || a[j].cols != m.cols) {
            for (double x : m.v) { r.diff_norm += x * x; r.max_abs = std::max(r.max_abs, std::abs(x)); }
            continue;
        }
        for (size_t i = 0; i < m.v.size(); ++i) {
            double d = a[j].v[i] - m.v[i];
            r.diff_norm += d * d; r.max_abs = std::max(r.max_abs, std::abs(d));
        }
    }
}
inline DiffReport compare(block_matrix const& a, block_matrix const& ref) { DiffReport r; accumulate_diff(a, ref, r); return r; }
inline DiffReport compare(Boundary const& a, Boundary const& ref)
{
    DiffReport r;
    if (a.aux_dim() != ref.aux_dim()) { r.structure_equal = 0; return r; }
    for (size_t b = 0; b < ref.aux_dim(); ++b) accumulate_diff(a[b], ref[b], r);
    return r;
}

// Dense effective Hamiltonian of a site problem: column i = site_hamil2(e_i). Returns eigenvalues ascending;
// asym = max |H - H^T|.
inline std::vector<double> dense_heff_spectrum(EngineIface& eng, MPSTensor const& templ, Boundary const& left, Boundary const& right,
                                               MPOTensor const& mpo, double* asym)
{
    templ.make_left_paired();
    size_t n = templ.data().num_elements();
    std::vector<double> H(n * n, 0.);
    auto flat_index = [&](block_matrix const& ref, Charge const& lc, Charge const& rc) -> long {
        long off = 0;
        for (size_t k = 0; k < ref.n_blocks(); ++k) {
            if (ref.basis().left_charge(k) == lc && ref.basis().right_charge(k) == rc) return off;
            off += (long)ref[k].v.size();
        }
        return -1;
    };
    size_t col = 0;
    for (size_t k = 0; k < templ.data().n_blocks(); ++k)
        for (size_t e = 0; e < templ.data()[k].v.size(); ++e, ++col) {
            MPSTensor x = templ;
            x.make_left_paired();
            x.data() *= 0.;
            x.data()[k].v[e] = 1.;
            MPSTensor y = eng.site_hamil2(x, left, right, mpo);
            y.make_left_paired();
            for (size_t kb = 0; kb < y.data().n_blocks(); ++kb) {
                long off = flat_index(templ.data(), y.data().basis().left_charge(kb), y.data().basis().right_charge(kb));
                if (off < 0) continue;
                Matrix const& yb = y.data()[kb];
                Matrix const& tb = templ.data()(y.data().basis().left_charge(kb), y.data().basis().right_charge(kb));
                for (size_t j = 0; j < std::min(yb.cols, tb.cols); ++j)
                    for (size_t i = 0; i < std::min(yb.rows, tb.rows); ++i) H[(off + i + j * tb.rows) + col * n] = yb(i, j);
            }
        }
    double a = 0;
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) a = std::max(a, std::abs(H[i + j * n] - H[j + i * n]));
    if (asym) *asym = a;
    for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) { double s = 0.5 * (H[i + j * n] + H[j + i * n]); H[i + j * n] = H[j + i * n] = s; }
    std::vector<double> w(n);
    int nn = (int)n, lwork = std::max(1, 3 * nn + 64), info = 0;
    std::vector<double> work(lwork);
    scipy_dsyev_("N", "U", &nn, H.data(), &nn, w.data(), work.data(), &lwork, &info);
    if (info != 0) throw std::runtime_error("dsyev failed");
    return w;
}

} // namespace qcm
