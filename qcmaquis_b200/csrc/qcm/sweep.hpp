// Host-side sweep driver above contraction::Engine, written against EngineIface so that the same code runs on the
// B200 engine and on the CPU checker.  It is the caller of the hot path, restated so that energies per micro-iteration
// can be compared engine against engine (north star: within 1e-8 Eh) and a sweep can be timed:
//   * single-site optimisation loop                  dmrg/optimize/ss_optimize.hpp:60-215  (noise alpha = 0: the site
//     tensor is re-orthogonalised by QR and the remainder is pushed into the neighbour, mpstensor.hpp normalize_left /
//     normalize_right, multiply_from_left / multiply_from_right)
//   * SiteProblem + ietl::mult                       dmrg/mp_tensors/siteproblem.h:22-53, optimize/ietl_lanczos_solver.h:108-115
//   * Jacobi-Davidson, ietl_jcd_gmres = 0            ietl/jacobi.h:361-451 (driver), :123-131 (correction t = -r + (r.u / u.u) u),
//     convergence test ietl/iteration.h (|r| <= max(rtol |theta|, atol)), defaults ietl_jcd_maxiter = 10, ietl_jcd_tol = 1e-8
//   * boundary bookkeeping                           optimize/optimize.h:105-165 (init_left_right, boundary_left/right_step)
// Everything dense here is small (block QR, a maxiter x maxiter eigenproblem); sigma and the boundary steps go through
// the engine.
#pragma once
#include "engine_iface.hpp"
#include "overlap.hpp"
#include <chrono>
#include <exception>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
void scipy_dgeqrf_(const int* m, const int* n, double* a, const int* lda, double* tau, double* work, const int* lwork, int* info);
void scipy_dorgqr_(const int* m, const int* n, const int* k, double* a, const int* lda, const double* tau, double* work, const int* lwork, int* info);
void scipy_dgelqf_(const int* m, const int* n, double* a, const int* lda, double* tau, double* work, const int* lwork, int* info);
void scipy_dorglq_(const int* m, const int* n, const int* k, double* a, const int* lda, const double* tau, double* work, const int* lwork, int* info);
void scipy_dsyev_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work, const int* lwork, int* info);
}

namespace qcm { namespace sweep {

// ---- block linear algebra (block_matrix_algorithms.h: gemm :48-70, qr / lq) ------------------------------------
// Blocks are independent: they are processed side by side on the host cores (one BLAS thread per block; run_blocks),
// the few blocks that are large enough to keep all cores busy on their own go first, one at a time, with threaded BLAS.
// Results are inserted in block order afterwards, so the outcome does not depend on the number of threads.
extern "C" void scipy_openblas_set_num_threads(int);
extern "C" int scipy_openblas_get_num_threads(void);
template <class Cost, class Body> inline void run_blocks(size_t n, Cost cost, Body body, double big = 2.5e10)
{
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
#ifdef _OPENMP
    const int nt = omp_get_max_threads();
    if (nt > 1 && !omp_in_parallel()) {
        const int blas_before = scipy_openblas_get_num_threads();
        size_t first_small = 0;
        while (first_small < n && cost(order[first_small]) >= big) ++first_small;
        if (first_small) { scipy_openblas_set_num_threads(nt); for (size_t q = 0; q < first_small; ++q) body(order[q]); }
        scipy_openblas_set_num_threads(1);
        std::exception_ptr err;
#pragma omp parallel for schedule(dynamic, 1)
        for (long q = (long)first_small; q < (long)n; ++q) {
            try { body(order[(size_t)q]); } catch (...) {
#pragma omp critical(qcm_run_blocks)
                err = std::current_exception();
            }
        }
        scipy_openblas_set_num_threads(blas_before);
        if (err) std::rethrow_exception(err);
        return;
    }
#endif
    for (size_t q = 0; q < n; ++q) body(order[q]);
}

inline void gemm(block_matrix const& A, block_matrix const& B, block_matrix& C)
{
    C.clear();
    struct Pair { size_t k, j; Matrix c; };
    std::vector<Pair> pairs;
    for (size_t k = 0; k < A.n_blocks(); ++k) {
        QnBlock const& a = A.basis()[k];
        for (size_t j = 0; j < B.n_blocks(); ++j) {
            QnBlock const& b = B.basis()[j];
            if (!(b.lc == a.rc)) continue;
            if (a.rs != b.ls) throw std::runtime_error("sweep::gemm: inner block sizes differ");
            pairs.push_back(Pair{k, j, Matrix()});
        }
    }
    run_blocks(pairs.size(), [&](size_t i) { return 2.0 * A.basis()[pairs[i].k].ls * A.basis()[pairs[i].k].rs * B.basis()[pairs[i].j].rs; },
               [&](size_t i) {
                   Pair& p = pairs[i];
                   p.c = Matrix(A.basis()[p.k].ls, B.basis()[p.j].rs);
                   dgemm(A.block(p.k), B.block(p.j), 1.0, 0.0, p.c.data(), p.c.rows);
               });
    for (Pair& p : pairs) C.match_and_add_block(p.c, A.basis()[p.k].lc, B.basis()[p.j].rc);
}
// A = Q R per block; Q: ls x k with orthonormal columns, R: k x rs, k = min(ls, rs)
inline void qr(block_matrix const& A, block_matrix& Q, block_matrix& R)
{
    Q.clear(); R.clear();
    const size_t nb = A.n_blocks();
    std::vector<Matrix> qs(nb), rs(nb);
    run_blocks(nb, [&](size_t b) { return 4.0 * A[b].rows * A[b].cols * std::min(A[b].rows, A[b].cols); }, [&](size_t b) {
        Matrix a = A[b];
        const int m = (int)a.rows, n = (int)a.cols, k = std::min(m, n);
        std::vector<double> tau(std::max(1, k)), work(1);
        int lwork = -1, info = 0;
        scipy_dgeqrf_(&m, &n, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        lwork = (int)work[0]; work.resize(std::max(1, lwork));
        scipy_dgeqrf_(&m, &n, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        if (info) throw std::runtime_error("dgeqrf failed");
        Matrix r(k, n);
        for (int j = 0; j < n; ++j) for (int i = 0; i <= std::min(j, k - 1); ++i) r(i, j) = a(i, j);
        lwork = -1;
        scipy_dorgqr_(&m, &k, &k, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        lwork = (int)work[0]; work.resize(std::max(1, lwork));
        scipy_dorgqr_(&m, &k, &k, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        if (info) throw std::runtime_error("dorgqr failed");
        Matrix q(m, k);
        for (int j = 0; j < k; ++j) for (int i = 0; i < m; ++i) q(i, j) = a(i, j);
        qs[b] = std::move(q); rs[b] = std::move(r);
    });
    for (size_t b = 0; b < nb; ++b) {
        Q.insert_block(qs[b], A.basis()[b].lc, A.basis()[b].rc);
        R.insert_block(rs[b], A.basis()[b].rc, A.basis()[b].rc);
    }
}
// A = L Q per block; L: ls x k, Q: k x rs with orthonormal rows
inline void lq(block_matrix const& A, block_matrix& L, block_matrix& Q)
{
    L.clear(); Q.clear();
    const size_t nb = A.n_blocks();
    std::vector<Matrix> ls(nb), qs(nb);
    run_blocks(nb, [&](size_t b) { return 4.0 * A[b].rows * A[b].cols * std::min(A[b].rows, A[b].cols); }, [&](size_t b) {
        Matrix a = A[b];
        const int m = (int)a.rows, n = (int)a.cols, k = std::min(m, n);
        std::vector<double> tau(std::max(1, k)), work(1);
        int lwork = -1, info = 0;
        scipy_dgelqf_(&m, &n, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        lwork = (int)work[0]; work.resize(std::max(1, lwork));
        scipy_dgelqf_(&m, &n, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        if (info) throw std::runtime_error("dgelqf failed");
        Matrix l(m, k);
        for (int j = 0; j < k; ++j) for (int i = j; i < m; ++i) l(i, j) = a(i, j);
        lwork = -1;
        scipy_dorglq_(&k, &n, &k, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        lwork = (int)work[0]; work.resize(std::max(1, lwork));
        scipy_dorglq_(&k, &n, &k, a.data(), &m, tau.data(), work.data(), &lwork, &info);
        if (info) throw std::runtime_error("dorglq failed");
        Matrix q(k, n);
        for (int j = 0; j < n; ++j) for (int i = 0; i < k; ++i) q(i, j) = a(i, j);
        ls[b] = std::move(l); qs[b] = std::move(q);
    });
    for (size_t b = 0; b < nb; ++b) {
        L.insert_block(ls[b], A.basis()[b].lc, A.basis()[b].lc);
        Q.insert_block(qs[b], A.basis()[b].lc, A.basis()[b].rc);
    }
}

// ---- MPSTensor normalisation (mpstensor.hpp normalize_left / normalize_right with the QR solver, multiply_from_*) ---
inline block_matrix normalize_left(MPSTensor& t)
{
    t.make_left_paired();
    block_matrix Q, R;
    qr(t.data(), Q, R);
    t.replace_left_paired(Q);
    return R;
}
inline block_matrix normalize_right(MPSTensor& t)
{
    t.make_right_paired();
    block_matrix L, Q;
    lq(t.data(), L, Q);
    t.replace_right_paired(Q);
    return L;
}
inline void multiply_from_left(MPSTensor& t, block_matrix const& N)
{
    t.make_right_paired();
    block_matrix tmp;
    gemm(N, t.data(), tmp);
    t.replace_right_paired(tmp);
}
inline void multiply_from_right(MPSTensor& t, block_matrix const& N)
{
    t.make_left_paired();
    block_matrix tmp;
    gemm(t.data(), N, tmp);
    t.replace_left_paired(tmp);
}
// MPS::canonize(0): right-normalise sites L-1 .. 1, then normalise site 0 (mps.hpp canonize / normalize_right)
inline void canonize_to_first(MPS& mps)
{
    for (size_t i = mps.size() - 1; i > 0; --i) {
        block_matrix l = normalize_right(mps[i]);
        multiply_from_right(mps[i - 1], l);
    }
    mps[0].divide_by_scalar(mps[0].scalar_norm());
}

// ---- Jacobi-Davidson on the site problem ---------------------------------------------------------------------
inline void axpy(MPSTensor& y, double a, MPSTensor const& x)       // y += a x, blocks matched by charge
{
    y.make_left_paired(); x.make_left_paired();
    y.data().axpy(a, x.data());
}
typedef EigenResult JDResult;

// contraction::site_ortho_boundaries (contractions/abelian/special.hpp:16-46): the component of an orthogonal state inside the
// variational space of the site being optimised,  o = ortho_left (x) ortho_mps (x) ortho_right^T, padded to the structure of mps.
// ortho_left / ortho_right are the MPS-MPS overlap matrices kept by the sweep (optimize.h:105-117).
inline MPSTensor site_ortho_boundaries(MPSTensor const& mps, MPSTensor const& ortho_mps, block_matrix const& ortho_left, block_matrix const& ortho_right)
{
    ortho_mps.make_right_paired();
    block_matrix t, t2, t3;
    gemm(ortho_left, ortho_mps.data(), t);
    reshape_right_to_left_new(mps.site_dim(), ortho_left.left_basis(), ortho_mps.col_dim(), t, t2);
    gemm(t2, transposed(ortho_right), t3);
    mps.make_left_paired();
    t = mps.data();
    reshape_and_pad_left(mps.site_dim(), ortho_left.left_basis(), ortho_right.left_basis(), mps.row_dim(), mps.col_dim(), t3, t);
    return MPSTensor(mps.site_dim(), mps.row_dim(), mps.col_dim(), t, LeftPaired);
}
// solve_ietl_jcd, optimize/ietl_jacobi_davidson.h:29-43: the local components of several orthogonal states are in general NOT
// orthogonal to each other, and the sequential projection of SingleSiteVS::project is exact only for mutually orthogonal vectors
// (without this step the solver's basis loses its orthonormality and the Ritz values stop being variational -- seen here as
// energies below the full-CI ground state with two orthogonal states).  Gram-Schmidt over the vectors in order; a vector whose
// remainder is shorter than thresholdForCompleteness = 1e-10 is neglected ("the corresponding boundary is too small"), the others
// are normalised.  (Two passes per vector instead of the reference's one: same span, better orthogonality.)
inline std::vector<MPSTensor> orthogonalised(std::vector<MPSTensor> vecs, double threshold = 1e-10)
{
    std::vector<MPSTensor> out;
    for (MPSTensor& v : vecs) {
        v.make_left_paired();
        for (int pass = 0; pass < 2; ++pass)
            for (MPSTensor const& o : out) axpy(v, -o.scalar_overlap(v), o);
        const double nrm = v.scalar_norm();
        if (nrm > threshold) { v.divide_by_scalar(nrm); out.push_back(v); }
    }
    return out;
}
// SingleSiteVS::project (optimize/ietl_lanczos_solver.h:90-95): t -= (o.t / o.o) o for every orthogonal vector
inline void project(MPSTensor& t, std::vector<MPSTensor> const& ortho_vecs)
{
    for (MPSTensor const& o : ortho_vecs) {
        const double oo = o.scalar_overlap(o);
        if (oo > 0.) axpy(t, -o.scalar_overlap(t) / oo, o);
    }
}

// ortho_vecs: states the solution is kept orthogonal to (excited states: the vector space projects them out at the three points
// of ietl/jacobi.h:378,393,432); with orthogonal states the recurrence runs on the host vectors, sigma through the engine
inline JDResult jacobi_davidson(EngineIface& eng, MPSTensor const& x0, Boundary const& left, Boundary const& right, MPOTensor const& mpo,
                                int max_iter, double tol, std::vector<MPSTensor> const& ortho_vecs_in = std::vector<MPSTensor>())
{
    JDResult res;
    if (ortho_vecs_in.empty() && eng.jacobi_davidson(x0, left, right, mpo, max_iter, tol, res)) return res;      // solver vectors kept on the device
    const std::vector<MPSTensor> ortho_vecs = orthogonalised(ortho_vecs_in);
    if (!ortho_vecs.empty()) {
        // ietl_jacobi_davidson.h:44-50,68-74: nothing left of the start vector outside the orthogonal states -- the space is
        // exhausted, the optimisation is skipped and the energy is the expectation value of the start vector
        MPSTensor tmp = x0; tmp.make_left_paired();
        for (MPSTensor const& o : ortho_vecs) axpy(tmp, -o.scalar_overlap(tmp), o);
        if (tmp.scalar_norm() < 1e-10) {
            MPSTensor hx = eng.site_hamil2(x0, left, right, mpo);
            hx.make_left_paired(); x0.make_left_paired();
            res.theta = x0.scalar_overlap(hx); res.vec = x0; res.n_sigma = 1;
            return res;
        }
    }
    std::vector<MPSTensor> V(max_iter + 1), VA(max_iter);
    std::vector<double> M((size_t)max_iter * max_iter, 0.);
    const double kappa = 0.25;
    V[0] = x0; V[0].make_left_paired();
    project(V[0], ortho_vecs);
    int it = 0;
    for (;;) {
        MPSTensor& t = V[it];
        // modified Gram-Schmidt with refinement
        const double tau = t.scalar_norm();
        for (int i = 0; i < it; ++i) axpy(t, -V[i].scalar_overlap(t), V[i]);
        if (t.scalar_norm() < kappa * tau)
            for (int i = 0; i < it; ++i) axpy(t, -V[i].scalar_overlap(t), V[i]);
        project(t, ortho_vecs);
        t.divide_by_scalar(t.scalar_norm());
        VA[it] = eng.site_hamil2(t, left, right, mpo);       // ietl::mult (y = H x; x.make_left_paired())
        VA[it].make_left_paired(); t.make_left_paired();
        res.n_sigma++;
        for (int i = 0; i <= it; ++i) M[(size_t)i + (size_t)it * max_iter] = V[i].scalar_overlap(VA[it]);
        // smallest eigenpair of the projected (it+1) x (it+1) matrix (upper triangle stored)
        const int dim = it + 1;
        std::vector<double> A((size_t)dim * dim), w(dim), work(std::max(1, 3 * dim));
        for (int c = 0; c < dim; ++c) for (int r = 0; r <= c; ++r) A[(size_t)r + (size_t)c * dim] = M[(size_t)r + (size_t)c * max_iter];
        int lwork = (int)work.size(), info = 0;
        scipy_dsyev_("V", "U", &dim, A.data(), &dim, w.data(), work.data(), &lwork, &info);
        if (info) throw std::runtime_error("dsyev failed in the Jacobi-Davidson subspace problem");
        const double theta = w[0];
        const double* s = A.data();
        MPSTensor u = V[0]; u.multiply_by_scalar(s[0]);
        for (int j = 1; j <= it; ++j) axpy(u, s[j], V[j]);
        MPSTensor r = VA[0]; r.multiply_by_scalar(s[0]);
        for (int j = 1; j <= it; ++j) axpy(r, s[j], VA[j]);
        project(r, ortho_vecs);
        axpy(r, -theta, u);
        ++it;
        const double rn = r.scalar_norm();
        res.theta = theta; res.resid = rn;
        if (rn <= tol * std::abs(theta) || rn <= tol || it >= max_iter) { res.vec = u; return res; }
        // correction without GMRES steps: t = -r + (r.u / u.u) u
        const double dru = r.scalar_overlap(u), duu = u.scalar_overlap(u);
        MPSTensor tn = r; tn.multiply_by_scalar(-1.);
        axpy(tn, dru / duu, u);
        V[it] = tn;
    }
}

// ---- single-site sweeps ------------------------------------------------------------------------------------------
struct SweepLog
{
    std::vector<double> energies;        // one per micro-iteration (site update): theta + core energy
    std::vector<double> sweep_energy;    // last energy of every sweep
    std::vector<double> sweep_seconds;
    std::vector<int> n_sigma;            // sigma evaluations per micro-iteration
    long total_sigma = 0;
    // two-site driver, wall seconds by phase: [0] two-site tensor (product, reshape, recoupling) [1] two-site MPO fusion
    // [2] eigensolver (sigma calls + host vector algebra) [3] split (reshape, SVD, normalisation) [4] boundary step
    double phase_seconds[5] = {0, 0, 0, 0, 0};
};

// Grow hook of the single-site loop (ss_optimize.hpp:168-195: mps.grow_l2r_sweep / grow_r2l_sweep replace the plain
// normalisation when the site has a neighbour in the sweep direction).  Returns false when it did nothing (alpha = 0 runs:
// the site tensor is re-orthogonalised by QR instead).  ts::NoiseGrow (twosite.hpp) is the noise-perturbed one.
struct NoGrow
{
    bool operator()(int /*lr*/, int /*site*/, MPS& /*mps*/, MPOTensor const& /*mpo*/, Boundary const& /*left*/, Boundary const& /*right*/) const { return false; }
};

// States the optimised state is kept orthogonal to (excited-state DMRG: optimize.h:44-75 ortho_mps, :105-117 the overlap
// boundaries ortho_left_ / ortho_right_ moved along with the sweep through Engine::overlap_left_step / overlap_right_step).
struct OrthoStates { std::vector<MPS> states; bool su2 = false; };

template <class Grow = NoGrow>
inline SweepLog ss_sweeps(EngineIface& eng, MPO const& mpo, MPS& mps, int nsweeps, int jcd_maxiter = 10, double jcd_tol = 1e-8, Grow grow = Grow(),
                          OrthoStates const* ortho = nullptr)
{
    const int L = (int)mps.size();
    SweepLog log;
    canonize_to_first(mps);
    std::vector<Boundary> left(L + 1), right(L + 1);
    left[0] = mps.left_boundary();
    right[L] = mps.right_boundary();
    for (int i = L - 1; i >= 0; --i) right[i] = eng.overlap_mpo_right_step(mps[i], mps[i], right[i + 1], mpo[i]);
    const int northo = ortho ? (int)ortho->states.size() : 0;
    std::vector<std::vector<block_matrix>> oleft((size_t)northo, std::vector<block_matrix>((size_t)L + 1)), oright = oleft;
    for (int n = 0; n < northo; ++n) {
        if ((int)ortho->states[(size_t)n].size() != L) throw std::runtime_error("ss_sweeps: orthogonal state of a different length");
        oleft[(size_t)n][0] = mps.left_boundary()[0];
        oright[(size_t)n][(size_t)L] = mps.right_boundary()[0];
        for (int i = L - 1; i >= 0; --i)
            oright[(size_t)n][(size_t)i] = overlap_right_step(eng, ortho->su2, mps[i], ortho->states[(size_t)n][(size_t)i], oright[(size_t)n][(size_t)i + 1]);
    }
    for (int sweep = 0; sweep < nsweeps; ++sweep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int _site = 0; _site < 2 * L; ++_site) {
            const int lr = _site < L ? +1 : -1;
            const int site = _site < L ? _site : 2 * L - _site - 1;
            std::vector<MPSTensor> ortho_vecs((size_t)northo);        // ss_optimize.hpp:107-111
            for (int n = 0; n < northo; ++n)
                ortho_vecs[(size_t)n] = site_ortho_boundaries(mps[site], ortho->states[(size_t)n][(size_t)site], oleft[(size_t)n][(size_t)site], oright[(size_t)n][(size_t)site + 1]);
            JDResult r = jacobi_davidson(eng, mps[site], left[site], right[site + 1], mpo[site], jcd_maxiter, jcd_tol, ortho_vecs);
            mps[site] = r.vec;
            log.energies.push_back(r.theta + mpo.core_energy);
            log.n_sigma.push_back(r.n_sigma); log.total_sigma += r.n_sigma;
            if (lr == +1) {
                if (!(site < L - 1 && grow(+1, site, mps, mpo[site], left[site], right[site + 1]))) {
                    block_matrix t = normalize_left(mps[site]);
                    if (site < L - 1) multiply_from_left(mps[site + 1], t);
                }
                left[site + 1] = eng.overlap_mpo_left_step(mps[site], mps[site], left[site], mpo[site]);
                for (int n = 0; n < northo; ++n)
                    oleft[(size_t)n][(size_t)site + 1] = overlap_left_step(eng, ortho->su2, mps[site], ortho->states[(size_t)n][(size_t)site], oleft[(size_t)n][(size_t)site]);
            } else {
                if (!(site > 0 && grow(-1, site, mps, mpo[site], left[site], right[site + 1]))) {
                    block_matrix t = normalize_right(mps[site]);
                    if (site > 0) multiply_from_right(mps[site - 1], t);
                }
                right[site] = eng.overlap_mpo_right_step(mps[site], mps[site], right[site + 1], mpo[site]);
                for (int n = 0; n < northo; ++n)
                    oright[(size_t)n][(size_t)site] = overlap_right_step(eng, ortho->su2, mps[site], ortho->states[(size_t)n][(size_t)site], oright[(size_t)n][(size_t)site + 1]);
            }
        }
        log.sweep_energy.push_back(log.energies.back());
        log.sweep_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    return log;
}

}} // namespace qcm::sweep
