// Host-side drivers around the hot path, written against EngineIface so that the same code runs on the B200
// engine and on the CPU checker: boundary chains (BoundaryPropagator.h:36-150, optimize.h:119-165), the
// sigma-energy identity of dmrg/tests/test_mps_mpo_ops/test_siteproblem.cpp:38-95, dense effective
// Hamiltonians for exact small cases, and comparison helpers.
#pragma once
#include "engine_iface.hpp"
#include "models.hpp"
#include "plan.hpp"
#include <chrono>

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


// ---------------------------------------------------------------------------------------------------------
// Synthetic mid-chain site problems for the large BASELINE configurations.  The MPO is the true
// one (built from the FCIDUMP by the term maker + TaggedMPOMaker restatement, so bond indexing, Hermitian
// pairs and bond spins are those the reference would produce); what is fabricated is the STATE: sector lists
// per bond from a smooth weight around the mean filling, truncated to a total bond dimension M, the boundary
// block structures obtained by propagating structure (no data) from both chain ends through the schedule
// builder, and N(0,1) contents.  Hermitian-skipped boundary entries stay empty, as in a real sweep.
inline std::vector<Index> synthetic_sectors(Problem const& P, size_t M)
{
    SymmKind symm = P.params.symm;
    int L = P.params.L;
    std::vector<Index> full = allowed_sectors(symm, P.model->lat.irreps, P.model->phys_indices, P.model->total_charge, (size_t)1 << 40);
    int nelec = is_su2(symm) ? P.params.nelec : P.params.nup + P.params.ndown;
    std::vector<Index> ret(L + 1);
    for (int i = 0; i <= L; ++i) {
        double nbar = (double)nelec * i / L;
        std::vector<double> w(full[i].size());
        double wsum = 0;
        for (size_t k = 0; k < full[i].size(); ++k) {
            Charge c = full[i][k].first;
            if (is_su2(symm)) {
                double dn = c[0] - nbar, s2 = c[1];
                w[k] = std::exp(-dn * dn / (2 * 1.3 * 1.3)) * (s2 + 1) * std::exp(-s2 * s2 / (2 * 2.0 * 2.0));
            } else {
                double du = c[0] - nbar / 2, dd = c[1] - nbar / 2;
                w[k] = std::exp(-(du * du + dd * dd) / (2 * 1.0 * 1.0));
            }
            wsum += w[k];
        }
        // proportional shares, capped by the sector's full dimension; leftover redistributed over uncapped sectors
        std::vector<double> share(w.size(), 0.);
        std::vector<char> capped(w.size(), 0);
        double remaining = (double)M;
        for (int pass = 0; pass < 8; ++pass) {
            double ws = 0;
            for (size_t k = 0; k < w.size(); ++k) if (!capped[k]) ws += w[k];
            if (ws <= 0) break;
            bool changed = false;
            for (size_t k = 0; k < w.size(); ++k) {
                if (capped[k]) continue;
                double want = remaining * w[k] / ws;
                if (want >= (double)full[i][k].second) { share[k] = (double)full[i][k].second; capped[k] = 1; changed = true; }
                else share[k] = want;
            }
            remaining = (double)M;
            for (size_t k = 0; k < w.size(); ++k) if (capped[k]) remaining -= share[k];
            if (!changed || remaining <= 0) break;
        }
        for (size_t k = 0; k < w.size(); ++k) {
            if (share[k] < 0.05) continue;                       // negligible weight: sector absent
            size_t sz = std::max<size_t>(1, (size_t)std::llround(share[k]));
            sz = std::min(sz, full[i][k].second);
            ret[i].insert(std::make_pair(full[i][k].first, sz));
        }
    }
    return ret;
}

struct SyntheticSite
{
    int site = 0; bool twosite = true;
    MPSTensor psi;
    Boundary left, right;
    MPOTensor const* mpo = nullptr;
    std::vector<Index> sectors;
    double setup_seconds = 0;
};

inline void fill_normal(block_matrix& m, uint64_t seed)
{
    std::mt19937_64 eng(seed);
    std::normal_distribution<double> nd(0., 1.);
    for (size_t k = 0; k < m.n_blocks(); ++k) for (auto& x : m[k].v) x = nd(eng);
}

inline Boundary boundary_from_layout(plan::BoundaryLayout const& L, uint64_t seed)
{
    Boundary b; b.resize(L.aux_dim());
    for (size_t k = 0; k < L.aux_dim(); ++k) {
        for (size_t j = 0; j < L.b[k].basis.size(); ++j) {
            QnBlock const& q = L.b[k].basis[j];
            b.raw(k).insert_block(Matrix(q.ls, q.rs), q.lc, q.rc);
        }
        fill_normal(b.raw(k), seed * 1000003ull + k);
    }
    return b;
}

inline SyntheticSite make_synthetic_site(Problem& P, int site, bool twosite, size_t M, unsigned seed)
{
    auto t0 = std::chrono::steady_clock::now();
    SyntheticSite S; S.site = site; S.twosite = twosite;
    SymmKind symm = P.symm();
    int L = P.params.L, last = twosite ? site + 1 : site;
    if (site < 0 || last >= L) throw std::runtime_error("make_synthetic_site: site outside the lattice");
    S.sectors = synthetic_sectors(P, M);
    auto zero = []() { return 0.0; };
    auto desc = [&](MPSTensor const& t) { t.make_left_paired(); return plan::TensorDesc{t.site_dim(), t.row_dim(), t.col_dim(), t.data().basis()}; };
    // structure chain from the left end
    std::vector<MPSTensor> tens(L);
    for (int i = 0; i < L; ++i) if (i < site || i > last) tens[i] = MPSTensor(P.phys(i), S.sectors[i], S.sectors[i + 1], zero);
    plan::BoundaryLayout ll, rl;
    {
        Index i0 = tens[0].left_i.size() ? tens[0].row_dim() : S.sectors[0];
        if (site == 0) i0 = S.sectors[0];
        std::vector<DualIndex> b0(1);
        for (auto const& e : i0) b0[0].insert(QnBlock(e.first, e.first, e.second, e.second));
        ll.assign(b0);
        for (int i = 0; i < site; ++i) {
            plan::Planner pl(symm, P.mpo[i], true);
            pl.structure_only = true;
            plan::TensorDesc d = desc(tens[i]);
            plan::Plan pp = pl.plan_left_step(d, d, ll);
            ll = pp.out_boundary;
        }
    }
    {
        Index iL = S.sectors[L];
        std::vector<DualIndex> bL(1);
        for (auto const& e : iL) bL[0].insert(QnBlock(e.first, e.first, e.second, e.second));
        rl.assign(bL);
        for (int i = L - 1; i > last; --i) {
            plan::Planner pl(symm, P.mpo[i], true);
            pl.structure_only = true;
            plan::TensorDesc d = desc(tens[i]);
            plan::Plan pp = pl.plan_right_step(d, d, rl);
            rl = pp.out_boundary;
        }
    }
    S.left = boundary_from_layout(ll, 7919ull * seed + 1);
    S.right = boundary_from_layout(rl, 7919ull * seed + 2);
    // the site tensor lives between the sector lists the neighbouring tensors actually kept
    Index li = site > 0 ? tens[site - 1].col_dim() : S.sectors[0];
    Index ri = last + 1 < L ? tens[last + 1].row_dim() : S.sectors[L];
    Index phys = twosite ? P.phys(site) * P.phys(site + 1) : P.phys(site);
    S.psi = MPSTensor(phys, li, ri, zero);
    fill_normal(S.psi.data(), 7919ull * seed + 3);
    S.psi.divide_by_scalar(S.psi.scalar_norm());
    S.mpo = twosite ? &P.twosite_mpo(site) : &P.mpo[site];
    S.setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return S;
}

// A random MPS on the synthetic sector lists (total bond dimension M at every bond, same lists as make_synthetic_site):
// the starting state of the sweep-level measurements at the large configurations.  Deterministic in (M, seed).
inline MPS make_synthetic_mps(Problem const& P, size_t M, unsigned seed)
{
    std::vector<Index> sectors = synthetic_sectors(P, M);
    const int L = P.params.L;
    MPS mps; mps.resize(L);
    std::mt19937_64 eng(1000003ull * seed + 17);
    std::uniform_real_distribution<double> ud(-1., 1.);
    for (int i = 0; i < L; ++i) {
        // the right index of tensor i must be what tensor i+1 keeps as its left index: build from the left with the
        // previous tensor's trimmed right index
        Index li = i == 0 ? sectors[0] : mps[i - 1].col_dim();
        mps[i] = MPSTensor(P.phys(i), li, sectors[i + 1], [&]() { return ud(eng); });
    }
    // trim dangling sectors from the right end (a sector of bond i+1 that tensor i+1 dropped cannot be populated)
    for (int i = L - 2; i >= 0; --i) {
        Index keep = mps[i + 1].row_dim();
        if (keep == mps[i].col_dim()) continue;
        mps[i].make_left_paired();
        block_matrix d;
        for (size_t k = 0; k < mps[i].data().n_blocks(); ++k) {
            Charge rc = mps[i].data().basis()[k].rc;
            if (keep.has(rc)) d.insert_block(mps[i].data()[k], mps[i].data().basis()[k].lc, rc);
        }
        mps[i] = MPSTensor(P.phys(i), mps[i].row_dim(), keep, d, LeftPaired);
    }
    return mps;
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
