// Two-site tensors: the data formats either side of the hot path in a two-site sweep, host code.
//   TwoSiteTensor            dmrg/mp_tensors/twositetensor.hpp:23-31 (product of two site tensors, both-paired),
//                            :119-139 (make_mps), :141-183 (split_mps_l2r / split_mps_r2l), :355-389 (operator<<)
//   reshape_both_to_right    dmrg/mp_tensors/ts_reshape.h:231-289      reshape_left_to_both  ts_reshape.h:89-144
//   reduce_right / unreduce_left (SU2: 6j recoupling of the two site spins)   dmrg/mp_tensors/ts_reduction.h:42-148,150-250
//   svd / estimate_truncation / svd_truncate   dmrg/block_matrix/block_matrix_algorithms.h:165-185,211-260,264-335
//   two-site sweep loop      dmrg/optimize/ts_optimize.hpp:60-270 (twosite_truncation = svd)
// The fused two-site physical index is phys_left * phys_right; the MPO side of the same fusion is make_twosite_mpo
// (mpo.hpp).  Everything here is O(tensor size) or block SVDs; sigma and the boundary steps go through the engine.
#pragma once
#include "sweep.hpp"
#include <cstring>
#include <functional>
#include "wigner.hpp"

extern "C" void scipy_dgesdd_(const char* jobz, const int* m, const int* n, double* a, const int* lda, double* s, double* u, const int* ldu,
                              double* vt, const int* ldvt, double* work, const int* lwork, int* iwork, int* info);

namespace qcm { namespace ts {

// ---- ts_reshape.h:231-289 --------------------------------------------------------------------------------------
inline void reshape_both_to_right(Index const& physical_i_left, Index const& physical_i_right, Index const& left_i, Index const& right_i,
                                  block_matrix const& m1, block_matrix& m2)
{
    m2 = block_matrix();
    Index phys2_i = physical_i_left * physical_i_right;
    ProductBasis phys_pb(physical_i_left, physical_i_right);
    ProductBasis in_left(physical_i_left, left_i);
    ProductBasis in_right(physical_i_right, right_i, true);
    ProductBasis out_right(phys2_i, right_i, true);
    for (size_t block = 0; block < m1.n_blocks(); ++block) {
        Charge blc = m1.basis()[block].lc, brc = m1.basis()[block].rc;
        Matrix const& in = m1[block];
        for (size_t s1 = 0; s1 < physical_i_left.size(); ++s1) {
            size_t l = left_i.position(fuse(blc, -physical_i_left[s1].first));
            if (l == left_i.size()) continue;
            for (size_t s2 = 0; s2 < physical_i_right.size(); ++s2) {
                size_t r = right_i.position(fuse(brc, physical_i_right[s2].first));
                if (r == right_i.size()) continue;
                Charge s_charge = fuse(physical_i_left[s1].first, physical_i_right[s2].first);
                Charge out_l_charge = left_i[l].first, out_r_charge = fuse(-s_charge, right_i[r].first);
                if (!m2.has_block(out_l_charge, out_r_charge))
                    m2.insert_block(Matrix(left_i[l].second, out_right.size(-s_charge, right_i[r].first)), out_l_charge, out_r_charge);
                Matrix& out = m2(out_l_charge, out_r_charge);
                // detail::reshape_b2r (alps_detail.hpp:112-126)
                const size_t in_left_offset = in_left(physical_i_left[s1].first, left_i[l].first), in_right_offset = in_right(physical_i_right[s2].first, right_i[r].first);
                const size_t out_right_offset = out_right(s_charge, right_i[r].first), out_phys_offset = phys_pb(physical_i_left[s1].first, physical_i_right[s2].first);
                const size_t sdim1 = physical_i_left[s1].second, sdim2 = physical_i_right[s2].second, ldim = left_i[l].second, rdim = right_i[r].second;
                for (size_t ss1 = 0; ss1 < sdim1; ++ss1)
                    for (size_t ss2 = 0; ss2 < sdim2; ++ss2) {
                        size_t ss_out = out_phys_offset + ss1 * sdim2 + ss2;
                        for (size_t rr = 0; rr < rdim; ++rr)
                            for (size_t ll = 0; ll < ldim; ++ll)
                                out(ll, out_right_offset + ss_out * rdim + rr) = in(in_left_offset + ss1 * ldim + ll, in_right_offset + ss2 * rdim + rr);
                    }
            }
        }
    }
}

// ---- ts_reshape.h:89-144 ---------------------------------------------------------------------------------------
inline void reshape_left_to_both(Index const& physical_i_left, Index const& physical_i_right, Index const& left_i, Index const& right_i,
                                 block_matrix const& m1, block_matrix& m2)
{
    m2 = block_matrix();
    Index phys2_i = physical_i_left * physical_i_right;
    ProductBasis phys_pb(physical_i_left, physical_i_right);
    ProductBasis in_left(phys2_i, left_i);
    ProductBasis out_right(physical_i_right, right_i, true);
    ProductBasis out_left(physical_i_left, left_i);
    for (size_t block = 0; block < m1.n_blocks(); ++block) {
        Charge blc = m1.basis()[block].lc, brc = m1.basis()[block].rc;
        Matrix const& in = m1[block];
        size_t r = right_i.position(brc);
        if (r == right_i.size()) throw std::runtime_error("m1 matrix inconsistent with right_i.");
        for (size_t s1 = 0; s1 < physical_i_left.size(); ++s1)
            for (size_t s2 = 0; s2 < physical_i_right.size(); ++s2) {
                Charge s_charge = fuse(physical_i_left[s1].first, physical_i_right[s2].first);
                size_t l = left_i.position(fuse(blc, -s_charge));
                if (l == left_i.size()) continue;
                Charge out_l_charge = fuse(physical_i_left[s1].first, left_i[l].first), out_r_charge = fuse(-physical_i_right[s2].first, right_i[r].first);
                size_t o = m2.find_block(out_l_charge, out_r_charge);
                if (o == m2.n_blocks())
                    o = m2.insert_block(Matrix(out_left.size(physical_i_left[s1].first, left_i[l].first), out_right.size(-physical_i_right[s2].first, right_i[r].first)),
                                        out_l_charge, out_r_charge);
                Matrix& out = m2[o];
                // detail::reshape_l2b (alps_detail.hpp:78-92)
                const size_t in_left_offset = in_left(s_charge, left_i[l].first), in_phys_offset = phys_pb(physical_i_left[s1].first, physical_i_right[s2].first);
                const size_t out_left_offset = out_left(physical_i_left[s1].first, left_i[l].first), out_right_offset = out_right(physical_i_right[s2].first, right_i[r].first);
                const size_t sdim1 = physical_i_left[s1].second, sdim2 = physical_i_right[s2].second, ldim = left_i[l].second, rdim = right_i[r].second;
                for (size_t ss1 = 0; ss1 < sdim1; ++ss1)
                    for (size_t ss2 = 0; ss2 < sdim2; ++ss2) {
                        size_t ss_out = in_phys_offset + ss1 * sdim2 + ss2;
                        for (size_t rr = 0; rr < rdim; ++rr)
                            for (size_t ll = 0; ll < ldim; ++ll)
                                out(out_left_offset + ss1 * ldim + ll, out_right_offset + ss2 * rdim + rr) = in(in_left_offset + ss_out * ldim + ll, rr);
                    }
            }
    }
}

// ---- ts_reduction.h:42-148: right-paired two-site tensor, uncoupled (s1, s2) columns -> spin-coupled columns --------
inline double recoupling(int jl, int jr, int j, int S2, int S1, int jm)
{
    double c = std::sqrt((j + 1.) * (jm + 1.)) * su2::wigner6j(jl, jr, j, S2, S1, jm);
    return (((jl + jr + S1 + S2) / 2) % 2) ? -c : c;
}
inline Index reduce_right(Index const& physical_i_left, Index const& physical_i_right, Index const& left_i, Index const& right_i,
                          block_matrix const& m1, block_matrix& m2)
{
    m2 = block_matrix();
    Index phys2_i = physical_i_left * physical_i_right;
    ProductBasis phys_pb(physical_i_left, physical_i_right);
    ProductBasis in_right(phys2_i, right_i, true);
    for (size_t block = 0; block < m1.n_blocks(); ++block) {
        Charge lc = m1.basis()[block].lc, in_r_charge = m1.basis()[block].rc;
        const size_t left_size = m1.basis()[block].ls;
        size_t o = m2.insert_block(Matrix(left_size, m1.basis()[block].rs), lc, in_r_charge);
        Matrix const& in_block = m1[block];
        Matrix& out_block = m2[o];
        auto reduce_r = [&](double scale, size_t in_right_offset, size_t in_phys_offset, size_t out_phys_offset, size_t sdim1, size_t sdim2, size_t ldim, size_t rdim) {
            for (size_t ss1 = 0; ss1 < sdim1; ++ss1)
                for (size_t ss2 = 0; ss2 < sdim2; ++ss2) {
                    size_t ss_in = in_phys_offset + ss1 * sdim2 + ss2, ss_out = out_phys_offset + ss1 * sdim2 + ss2;
                    const double* src = &in_block(0, in_right_offset + ss_in * rdim);
                    double* dst = &out_block(0, in_right_offset + ss_out * rdim);
                    for (size_t i = 0; i < ldim * rdim; ++i) dst[i] += scale * src[i];
                }
        };
        for (size_t s = 0; s < phys2_i.size(); ++s) {
            Charge s_charge = phys2_i[s].first;
            size_t r = right_i.position(fuse(in_r_charge, s_charge));
            if (r == right_i.size()) continue;
            size_t in_right_offset = in_right(s_charge, right_i[r].first);
            size_t right_size = right_i[r].second;
            for (size_t s1 = 0; s1 < physical_i_left.size(); ++s1)
                for (size_t s2 = 0; s2 < physical_i_right.size(); ++s2) {
                    Charge phys_c1 = physical_i_left[s1].first, phys_c2 = physical_i_right[s2].first;
                    if (!(s_charge == fuse(phys_c1, phys_c2))) continue;
                    size_t in_phys_offset = phys_pb(phys_c1, phys_c2);
                    int S1 = std::abs(spin(phys_c1)), S2 = std::abs(spin(phys_c2));
                    int jl = spin(lc), jm = spin(lc) + spin(phys_c1), jr = spin(right_i[r].first);
                    if (jm < 0) continue;
                    if (jl == jr && jl > 0 && S1 == 1 && S2 == 1) {
                        size_t base_offset = (spin(phys_c1) == 1) ? in_phys_offset : in_phys_offset - 1;
                        for (int j = std::abs(S1 - S2); j <= std::abs(S1 + S2); j += 2) {
                            size_t out_phys_offset = base_offset + j / 2;
                            reduce_r(recoupling(jl, jr, j, S2, S1, jm), in_right_offset, in_phys_offset, out_phys_offset,
                                     physical_i_left[s1].second, physical_i_right[s2].second, left_size, right_size);
                        }
                    } else {
                        int j = std::abs(spin(phys_c1) + spin(phys_c2));
                        reduce_r(recoupling(jl, jr, j, S2, S1, jm), in_right_offset, in_phys_offset, in_phys_offset,
                                 physical_i_left[s1].second, physical_i_right[s2].second, left_size, right_size);
                    }
                }
        }
    }
    return phys2_i;
}

// ---- ts_reduction.h:150-250: left-paired two-site tensor, spin-coupled rows -> uncoupled (s1, s2) rows -----------------
inline Index unreduce_left(Index const& physical_i_left, Index const& physical_i_right, Index const& left_i, Index const& right_i,
                           block_matrix const& m1, block_matrix& m2)
{
    m2 = block_matrix();
    Index phys2_i = physical_i_left * physical_i_right;
    ProductBasis phys_pb(physical_i_left, physical_i_right);
    ProductBasis in_left(phys2_i, left_i);
    for (size_t block = 0; block < m1.n_blocks(); ++block) {
        Charge rc = m1.basis()[block].rc, in_l_charge = m1.basis()[block].lc;
        const size_t right_size = m1.basis()[block].rs;
        size_t o = m2.insert_block(Matrix(m1.basis()[block].ls, right_size), in_l_charge, rc);
        Matrix const& in_block = m1[block];
        Matrix& out_block = m2[o];
        auto reduce_l = [&](double scale, size_t in_left_offset, size_t in_phys_offset, size_t out_phys_offset, size_t sdim1, size_t sdim2, size_t ldim, size_t rdim) {
            for (size_t ss1 = 0; ss1 < sdim1; ++ss1)
                for (size_t ss2 = 0; ss2 < sdim2; ++ss2) {
                    size_t ss_in = in_phys_offset + ss1 * sdim2 + ss2, ss_out = out_phys_offset + ss1 * sdim2 + ss2;
                    for (size_t rr = 0; rr < rdim; ++rr)
                        for (size_t ll = 0; ll < ldim; ++ll)
                            out_block(in_left_offset + ss_out * ldim + ll, rr) += scale * in_block(in_left_offset + ss_in * ldim + ll, rr);
                }
        };
        for (size_t s = 0; s < phys2_i.size(); ++s) {
            Charge s_charge = phys2_i[s].first;
            size_t l = left_i.position(fuse(in_l_charge, -s_charge));
            if (l == left_i.size()) continue;
            size_t in_left_offset = in_left(s_charge, left_i[l].first);
            for (size_t s1 = 0; s1 < physical_i_left.size(); ++s1)
                for (size_t s2 = 0; s2 < physical_i_right.size(); ++s2) {
                    Charge phys_c1 = physical_i_left[s1].first, phys_c2 = physical_i_right[s2].first;
                    if (!(s_charge == fuse(phys_c1, phys_c2))) continue;
                    size_t in_phys_offset = phys_pb(phys_c1, phys_c2);
                    int S1 = std::abs(spin(phys_c1)), S2 = std::abs(spin(phys_c2));
                    int jl = spin(left_i[l].first), jr = spin(rc);
                    if (jl == jr && jl > 0 && S1 == 1 && S2 == 1) {
                        int j = (spin(phys_c1) == 1) ? 0 : 2;
                        size_t base_offset = (j == 0) ? in_phys_offset : in_phys_offset - 1;
                        for (int jm = jl - 1; jm <= jl + 1; jm += 2) {
                            size_t out_phys_offset = (jm == jl - 1) ? base_offset + 1 : base_offset;
                            reduce_l(recoupling(jl, jr, j, S2, S1, jm), in_left_offset, in_phys_offset, out_phys_offset,
                                     physical_i_left[s1].second, physical_i_right[s2].second, left_i[l].second, right_size);
                        }
                    } else {
                        int j = std::abs(spin(phys_c1) + spin(phys_c2));
                        int jm = jl + spin(phys_c1);
                        if (jm < 0) continue;
                        reduce_l(recoupling(jl, jr, j, S2, S1, jm), in_left_offset, in_phys_offset, in_phys_offset,
                                 physical_i_left[s1].second, physical_i_right[s2].second, left_i[l].second, right_size);
                    }
                }
        }
    }
    return phys2_i;
}

// ---- block SVD with truncation (block_matrix_algorithms.h:165-185,211-260,264-335) ----------------------------------
struct Truncation { size_t bond_dimension = 0; double truncated_weight = 0, truncated_fraction = 0, smallest_ev = 0; };
// wall seconds inside the split, by part (development aid, printed by the drivers under QCM_DEBUG):
// [0] block SVDs [1] combination across ranks [2] truncation + tensor assembly [3] reshapes / recoupling [4] normalisation + shift
inline double* split_seconds() { static double s[5] = {0, 0, 0, 0, 0}; return s; }
struct SplitClock
{
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(int i) { auto n = std::chrono::steady_clock::now(); split_seconds()[i] += std::chrono::duration<double>(n - t).count(); t = n; }
};

// With many blocks the SVDs run side by side, one BLAS thread each, largest first (a threaded dgesdd gains a factor of two on
// eight cores and blocks everything else); only when there are fewer blocks than cores do the large ones get threaded BLAS.
inline double svd_threaded_threshold(size_t n_blocks)
{
#ifdef _OPENMP
    if (n_blocks * 2 >= (size_t)omp_get_max_threads()) return 1e300;
#endif
    (void)n_blocks;
    return 2.5e10;
}
// M = U diag(S) V per block; singular values below max(rel_tol * largest, the (Mmax+1)-th largest) are dropped.
// With several ranks (eng->comm_world() > 1) the blocks are divided among the ranks, largest first to the least loaded one,
// and the factors are combined with one allreduce of a zero-padded buffer: the split costs 1/world of the host time per
// rank and every rank holds bit-identical factors, whatever its BLAS threads did.
inline Truncation svd_truncate(block_matrix const& M, block_matrix& U, block_matrix& V, std::vector<std::vector<double>>& S, double rel_tol, size_t Mmax,
                               EngineIface* eng = nullptr)
{
    const size_t nb = M.n_blocks();
    std::vector<Matrix> us(nb), vs(nb);
    S.assign(nb, std::vector<double>());
    const int world = eng ? eng->comm_world() : 1, rank = eng ? eng->comm_rank() : 0;
    auto cost = [&](size_t b) { double m = (double)M[b].rows, n = (double)M[b].cols; return 20.0 * m * n * std::min(m, n); };
    SplitClock clk;
    std::vector<int> owner(nb, 0);
    std::vector<size_t> mine;
    {
        std::vector<size_t> order(nb);
        std::iota(order.begin(), order.end(), (size_t)0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost(a) > cost(b); });
        std::vector<double> load((size_t)world, 0.);
        for (size_t b : order) { int r = (int)(std::min_element(load.begin(), load.end()) - load.begin()); owner[b] = r; load[(size_t)r] += cost(b); if (r == rank) mine.push_back(b); }
    }
    sweep::run_blocks(mine.size(), [&](size_t q) { return cost(mine[q]); }, [&](size_t q) {
        const size_t b = mine[q];
        Matrix a = M[b];
        const int m = (int)a.rows, n = (int)a.cols, k = std::min(m, n);
        us[b] = Matrix(m, k); vs[b] = Matrix(k, n); S[b].assign(k, 0.);
        std::vector<double> work(1); std::vector<int> iwork(8 * std::max(1, k));
        int lwork = -1, info = 0;
        scipy_dgesdd_("S", &m, &n, a.data(), &m, S[b].data(), us[b].data(), &m, vs[b].data(), &k, work.data(), &lwork, iwork.data(), &info);
        lwork = (int)work[0]; work.resize(std::max(1, lwork));
        scipy_dgesdd_("S", &m, &n, a.data(), &m, S[b].data(), us[b].data(), &m, vs[b].data(), &k, work.data(), &lwork, iwork.data(), &info);
        if (info) throw std::runtime_error("dgesdd failed");
    }, svd_threaded_threshold(mine.size()));
    clk.lap(0);
    if (world > 1) {
        std::vector<size_t> off(nb + 1, 0);
        for (size_t b = 0; b < nb; ++b) { size_t m = M[b].rows, n = M[b].cols, k = std::min(m, n); off[b + 1] = off[b] + m * k + k + k * n; }
        std::vector<double> buf(off[nb], 0.);
        for (size_t b : mine) {
            double* p = buf.data() + off[b];
            p = std::copy(us[b].v.begin(), us[b].v.end(), p); p = std::copy(S[b].begin(), S[b].end(), p); std::copy(vs[b].v.begin(), vs[b].v.end(), p);
        }
        eng->allreduce_sum(buf.data(), buf.size());
        for (size_t b = 0; b < nb; ++b) {
            const size_t m = M[b].rows, n = M[b].cols, k = std::min(m, n);
            const double* p = buf.data() + off[b];
            us[b] = Matrix(m, k); vs[b] = Matrix(k, n); S[b].assign(k, 0.);
            std::copy(p, p + m * k, us[b].v.begin()); p += m * k;
            std::copy(p, p + k, S[b].begin()); p += k;
            std::copy(p, p + k * n, vs[b].v.begin());
        }
        clk.lap(1);
    }
    // estimate_truncation
    std::vector<double> all;
    for (auto const& s : S) all.insert(all.end(), s.begin(), s.end());
    if (all.empty()) throw std::runtime_error("svd_truncate: empty matrix");
    std::sort(all.begin(), all.end(), std::greater<double>());
    double cut = rel_tol * all[0];
    if (all.size() > Mmax) cut = std::max(cut, all[Mmax]);
    Truncation tr;
    tr.smallest_ev = cut / all[0];
    double sum1 = 0, sum2 = 0;
    for (double x : all) { sum1 += x; sum2 += x * x; if (x < cut) { tr.truncated_fraction += x; tr.truncated_weight += x * x; } }
    tr.truncated_fraction /= sum1; tr.truncated_weight /= sum2;
    U.clear(); V.clear();
    std::vector<std::vector<double>> Skept;
    for (size_t b = 0; b < nb; ++b) {
        size_t keep = std::find_if(S[b].begin(), S[b].end(), [cut](double x) { return x < cut; }) - S[b].begin();
        if (keep == 0) continue;
        Matrix u = us[b], v = vs[b];
        u.resize(u.rows, keep); v.resize(keep, v.cols);
        Charge lc = M.basis()[b].lc, rc = M.basis()[b].rc;
        size_t iu = U.insert_block(u, lc, rc);
        V.insert_block(v, lc, rc);                  // the new bond carries the block's charges on both factors (m = M.left_basis())
        Skept.insert(Skept.begin() + iu, std::vector<double>(S[b].begin(), S[b].begin() + keep));
        tr.bond_dimension += keep;
    }
    S.swap(Skept);
    clk.lap(2);
    return tr;
}
// ---- heev_truncate (block_matrix_algorithms.h:187-200,211-262,387-440): eigen-decomposition of the (perturbed) reduced density
// matrix per block, eigenvalues descending; eigenvalues below max(cutoff * largest, the (Mmax+1)-th largest) are dropped.
// Eigenvector signs: largest-magnitude component positive (adjustPhase).
inline Truncation heev_truncate(block_matrix const& M, block_matrix& evecs, std::vector<std::vector<double>>& evals, double cutoff, size_t Mmax)
{
    const size_t nb = M.n_blocks();
    std::vector<Matrix> vs(nb);
    evals.assign(nb, std::vector<double>());
    sweep::run_blocks(nb, [&](size_t b) { double n = (double)M[b].rows; return 10.0 * n * n * n; }, [&](size_t b) {
        Matrix a = M[b];
        const int n = (int)a.rows;
        if (a.rows != a.cols) throw std::runtime_error("heev_truncate: density-matrix block is not square");
        std::vector<double> w((size_t)n), work(1);
        int lwork = -1, info = 0;
        scipy_dsyev_("V", "U", &n, a.data(), &n, w.data(), work.data(), &lwork, &info);
        lwork = (int)work[0]; work.resize(std::max(1, lwork));
        scipy_dsyev_("V", "U", &n, a.data(), &n, w.data(), work.data(), &lwork, &info);
        if (info) throw std::runtime_error("dsyev failed");
        Matrix v((size_t)n, (size_t)n);
        evals[b].resize((size_t)n);
        for (int j = 0; j < n; ++j) {          // descending
            evals[b][(size_t)j] = w[(size_t)(n - 1 - j)];
            size_t imax = 0;
            for (int i = 0; i < n; ++i) { v((size_t)i, (size_t)j) = a((size_t)i, (size_t)(n - 1 - j)); if (std::abs(v((size_t)i, (size_t)j)) > std::abs(v(imax, (size_t)j))) imax = (size_t)i; }
            if (v(imax, (size_t)j) < 0) for (int i = 0; i < n; ++i) v((size_t)i, (size_t)j) = -v((size_t)i, (size_t)j);
        }
        vs[b] = std::move(v);
    });
    std::vector<double> all;
    for (auto const& e : evals) all.insert(all.end(), e.begin(), e.end());
    if (all.empty()) throw std::runtime_error("heev_truncate: empty matrix");
    std::sort(all.begin(), all.end(), std::greater<double>());
    double cut = cutoff * all[0];
    if (all.size() > Mmax) cut = std::max(cut, all[Mmax]);
    Truncation tr;
    tr.smallest_ev = cut / all[0];
    double sum1 = 0;
    for (double x : all) { sum1 += x; if (x < cut) tr.truncated_fraction += x; }
    tr.truncated_fraction /= sum1; tr.truncated_weight = tr.truncated_fraction;
    evecs.clear();
    std::vector<std::vector<double>> kept;
    for (size_t b = 0; b < nb; ++b) {
        size_t keep = std::find_if(evals[b].begin(), evals[b].end(), [cut](double x) { return x < cut; }) - evals[b].begin();
        if (keep == 0) continue;
        Matrix v = vs[b];
        v.resize(v.rows, keep);
        size_t iu = evecs.insert_block(v, M.basis()[b].lc, M.basis()[b].rc);
        kept.insert(kept.begin() + iu, std::vector<double>(evals[b].begin(), evals[b].begin() + keep));
        tr.bond_dimension += keep;
    }
    evals.swap(kept);
    return tr;
}
using qcm::transposed;
// the blocks of a noise term its consumer keeps: those the tensor's own block structure has (SU2: Y Y^T also couples
// different spin sectors through a common column sector; these never enter the density matrix)
inline block_matrix noise_kept(block_matrix const& noise, DualIndex const& keep_basis)
{
    block_matrix r;
    for (size_t k = 0; k < noise.n_blocks(); ++k)
        if (keep_basis.has(noise.basis()[k].lc, noise.basis()[k].rc)) r.insert_block(noise[k], noise.basis()[k].lc, noise.basis()[k].rc);
    return r;
}
// dm += alpha * noise, only where dm has a block (twositetensor.hpp:204-219; prediction.hpp:38-46)
inline void add_noise(block_matrix& dm, block_matrix const& noise, double alpha, DualIndex const& keep_basis)
{
    for (size_t k = 0; k < noise.n_blocks(); ++k) {
        Charge lc = noise.basis()[k].lc, rc = noise.basis()[k].rc;
        if (!keep_basis.has(lc, rc)) continue;
        Matrix t = noise[k];
        for (double& x : t.v) x *= alpha;
        dm.match_and_add_block(t, lc, rc);
    }
}

// ---- single-site subspace expansion: MPS::grow_l2r_sweep / grow_r2l_sweep (mps.hpp:213-240) =
// predict_new_state_*_sweep (prediction.hpp:19-57, 84-130; doPerturbDM) followed by predict_lanczos_*_sweep (:59-82, 132-156).
// The reduced density matrix of the optimised site tensor is perturbed by alpha times the engine's noise term, its leading
// eigenvectors become the new site tensor U, and U^T M (M U^T) is pushed into the neighbour, whose bond grows accordingly.
struct NoiseGrow
{
    EngineIface& eng;
    double alpha, cutoff;
    size_t Mmax;
    std::vector<Truncation>* log = nullptr;
    bool operator()(int lr, int site, MPS& mps, MPOTensor const& mpo, Boundary const& left, Boundary const& right) const
    {
        MPSTensor& cur = mps[(size_t)site];
        block_matrix dm, U; std::vector<std::vector<double>> S;
        Truncation tr;
        if (lr == +1) {
            cur.make_left_paired();
            sweep::gemm(cur.data(), transposed(cur.data()), dm);
            block_matrix nz = eng.noise_left(cur, left, mpo);
            cur.make_left_paired();
            add_noise(dm, nz, alpha, cur.data().basis());
            tr = heev_truncate(dm, U, S, cutoff, Mmax);
            block_matrix rest;
            sweep::gemm(transposed(U), cur.data(), rest);            // getZeroSiteTensorL2R: U^T M
            MPSTensor& next = mps[(size_t)site + 1];
            sweep::multiply_from_left(next, rest);
            cur.replace_left_paired(U);
        } else {
            cur.make_right_paired();
            sweep::gemm(transposed(cur.data()), cur.data(), dm);
            block_matrix nz = eng.noise_right(cur, right, mpo);
            cur.make_right_paired();                                 // engines may re-pair their operand (prediction.hpp:115)
            add_noise(dm, nz, alpha, cur.data().basis());
            tr = heev_truncate(dm, U, S, cutoff, Mmax);
            block_matrix rest;
            sweep::gemm(cur.data(), U, rest);                        // getZeroSiteTensorR2L: M U  (U holds adjoint(U^T)'s columns)
            MPSTensor& prev = mps[(size_t)site - 1];
            sweep::multiply_from_right(prev, rest);
            cur.replace_right_paired(transposed(U));
        }
        if (log) log->push_back(tr);
        return true;
    }
};

// diag(S) * V and U * diag(S), block by block (U, V, S in the same block order)
inline block_matrix scale_rows(block_matrix const& V, std::vector<std::vector<double>> const& S)
{
    block_matrix r = V;
    for (size_t b = 0; b < r.n_blocks(); ++b) for (size_t j = 0; j < r[b].cols; ++j) for (size_t i = 0; i < r[b].rows; ++i) r[b](i, j) *= S[b][i];
    return r;
}
inline block_matrix scale_cols(block_matrix const& U, std::vector<std::vector<double>> const& S)
{
    block_matrix r = U;
    for (size_t b = 0; b < r.n_blocks(); ++b) for (size_t j = 0; j < r[b].cols; ++j) for (size_t i = 0; i < r[b].rows; ++i) r[b](i, j) *= S[b][j];
    return r;
}

// ---- twositetensor.hpp ----------------------------------------------------------------------------------------------
class TwoSiteTensor
{
public:
    TwoSiteTensor(SymmKind symm, MPSTensor const& mps1, MPSTensor const& mps2)
        : su2_(is_su2(symm)), phys_i(mps1.site_dim() * mps2.site_dim()), phys_i_left(mps1.site_dim()), phys_i_right(mps2.site_dim()),
          left_i(mps1.row_dim()), right_i(mps2.col_dim()), storage_(Both)
    {
        mps1.make_left_paired(); mps2.make_right_paired();
        sweep::gemm(mps1.data(), mps2.data(), data_);
    }
    // :119-139 -- the tensor the site problem works on (right-paired; SU2: spin-coupled two-site basis)
    MPSTensor make_mps() const
    {
        block_matrix rp;
        if (storage_ != Both) throw std::runtime_error("TwoSiteTensor::make_mps: tensor is not both-paired");
        reshape_both_to_right(phys_i_left, phys_i_right, left_i, right_i, data_, rp);
        if (!su2_) return MPSTensor(phys_i, left_i, right_i, rp, RightPaired);
        block_matrix tmp;
        Index phys_out = reduce_right(phys_i_left, phys_i_right, left_i, right_i, rp, tmp);
        return MPSTensor(phys_out, left_i, right_i, tmp, RightPaired);
    }
    // :355-389 -- take the optimised tensor back (SU2: undo the spin coupling)
    TwoSiteTensor& operator<<(MPSTensor const& rhs)
    {
        rhs.make_left_paired();
        if (su2_) {
            block_matrix tmp;
            phys_i = unreduce_left(phys_i_left, phys_i_right, left_i, right_i, rhs.data(), tmp);
            data_ = tmp;
        } else
            data_ = rhs.data();
        left_i = rhs.row_dim(); right_i = rhs.col_dim();
        storage_ = Left;
        return *this;
    }
    // :141-183
    void split_mps_l2r(size_t Mmax, double cutoff, MPSTensor& t1, MPSTensor& t2, Truncation& trunc, EngineIface* eng = nullptr)
    {
        make_both_paired();
        block_matrix u, v; std::vector<std::vector<double>> s;
        trunc = svd_truncate(data_, u, v, s, cutoff, Mmax, eng);
        t1 = MPSTensor(phys_i_left, left_i, u.right_basis(), u, LeftPaired);
        block_matrix sv = scale_rows(v, s);
        t2 = MPSTensor(phys_i_right, sv.left_basis(), right_i, sv, RightPaired);
    }
    void split_mps_r2l(size_t Mmax, double cutoff, MPSTensor& t1, MPSTensor& t2, Truncation& trunc, EngineIface* eng = nullptr)
    {
        make_both_paired();
        block_matrix u, v; std::vector<std::vector<double>> s;
        trunc = svd_truncate(data_, u, v, s, cutoff, Mmax, eng);
        t2 = MPSTensor(phys_i_right, v.left_basis(), right_i, v, RightPaired);
        block_matrix us = scale_cols(u, s);
        t1 = MPSTensor(phys_i_left, left_i, us.right_basis(), us, LeftPaired);
    }

    // the both-paired two-site data seen as a tensor of site 1 with a fat right index / of site 2 with a fat left index: the
    // operands predict_split_l2r / r2l hand to left / right_boundary_tensor_mpo (:195-196, :247-248)
    MPSTensor fat_right_tensor()
    {
        make_both_paired();
        return MPSTensor(phys_i_left, left_i, adjoin(phys_i_right) * right_i, data_, LeftPaired);
    }
    DualIndex both_paired_basis() { make_both_paired(); return data_.basis(); }
    MPSTensor fat_left_tensor()
    {
        make_both_paired();
        return MPSTensor(phys_i_right, phys_i_left * left_i, right_i, data_, RightPaired);
    }
    // :184-233 -- split with the noise-perturbed reduced density matrix (left index open): dm = T T^T + alpha * sum_b Y_b Y_b^T with
    // Y = left_boundary_tensor_mpo of the two-site tensor seen as a site-1 tensor with a fat right index; heev_truncate
    void predict_split_l2r(size_t Mmax, double cutoff, double alpha, Boundary const& left, MPOTensor const& mpo_site1, EngineIface& eng,
                           MPSTensor& t1, MPSTensor& t2, Truncation& trunc)
    {
        make_both_paired();
        block_matrix dm;
        sweep::gemm(data_, transposed(data_), dm);
        if (alpha != 0.) {
            add_noise(dm, eng.noise_left(fat_right_tensor(), left, mpo_site1), alpha, data_.basis());
        }
        block_matrix U; std::vector<std::vector<double>> S;
        trunc = heev_truncate(dm, U, S, cutoff, Mmax);
        t1 = MPSTensor(phys_i_left, left_i, U.right_basis(), U, LeftPaired);
        block_matrix V;
        sweep::gemm(transposed(U), data_, V);
        t2 = MPSTensor(phys_i_right, V.left_basis(), right_i, V, RightPaired);
    }
    // :236-300 -- right index open
    void predict_split_r2l(size_t Mmax, double cutoff, double alpha, Boundary const& right, MPOTensor const& mpo_site2, EngineIface& eng,
                           MPSTensor& t1, MPSTensor& t2, Truncation& trunc)
    {
        make_both_paired();
        block_matrix dm;
        sweep::gemm(transposed(data_), data_, dm);
        if (alpha != 0.) {
            add_noise(dm, eng.noise_right(fat_left_tensor(), right, mpo_site2), alpha, data_.basis());
        }
        block_matrix U; std::vector<std::vector<double>> S;
        trunc = heev_truncate(dm, U, S, cutoff, Mmax);
        block_matrix Ut = transposed(U);
        t2 = MPSTensor(phys_i_right, Ut.left_basis(), right_i, Ut, RightPaired);
        block_matrix V;
        sweep::gemm(data_, U, V);
        t1 = MPSTensor(phys_i_left, left_i, V.right_basis(), V, LeftPaired);
    }

private:
    enum Storage { Both, Left };
    void make_both_paired()
    {
        if (storage_ == Both) return;
        block_matrix tmp;
        reshape_left_to_both(phys_i_left, phys_i_right, left_i, right_i, data_, tmp);
        data_ = tmp; storage_ = Both;
    }
    bool su2_;
    Index phys_i, phys_i_left, phys_i_right, left_i, right_i;
    block_matrix data_;
    Storage storage_;
};

// ---- two-site sweeps (ts_optimize.hpp:60-270, twosite_truncation = svd) ------------------------------------------------
struct TsParams
{
    size_t Mmax = 100; double cutoff = 1e-16; int jcd_maxiter = 10; double jcd_tol = 1e-8;
    // A boundary that the sweep has moved past is recomputed on the way back before it is read again
    // (optimize.h:150-165): with drop_stale its storage is released right away, so that at most L + 1 boundaries are
    // resident at any time (the reference's storage::drop on the disk tier, utils/storage.h:176-181).
    bool drop_stale = false;
    // noise parameter of the perturbed truncation (ts_optimize.hpp:198-215 predict_split_l2r / predict_split_r2l; the
    // reference's default schedule starts at alpha_initial = 1e-2, DmrgParameters.h:47-49).  0: plain SVD split
    double alpha = 0.;
    // states the optimised state is kept orthogonal to (excited states; ts_optimize.hpp:120-128, optimize.h:105-117)
    sweep::OrthoStates const* ortho = nullptr;
    // storage protocol of ts_optimize.hpp:92-118: the boundaries of the next site are prefetched while this site is solved,
    // the one the sweep leaves behind is evicted (it is needed again only on the way back); with spill the engine keeps
    // three to four boundaries in its fast tier instead of L + 1
    bool spill = false;
    // stop after this many micro-iterations of the LAST sweep (0: full sweeps); measurement aid for bounded runs
    int max_micro_iterations = 0;
    // asked before every micro-iteration; true ends the run at that site boundary (wall-clock budgets of measurement
    // runs; with several ranks the callee must return the same answer on all of them)
    std::function<bool()> should_stop;
};

template <class TsMpo>      // TsMpo(p) -> MPOTensor const& of the fused sites (p, p+1)  (ts_ops.h make_ts_cache_mpo)
inline sweep::SweepLog ts_sweeps(SymmKind symm, EngineIface& eng, MPO const& mpo, TsMpo ts_mpo, MPS& mps, int nsweeps, TsParams const& prm,
                                 std::vector<size_t>* bond_dims = nullptr, double* init_seconds = nullptr)
{
    const int L = (int)mps.size();
    sweep::SweepLog log;
    auto t_init = std::chrono::steady_clock::now();
    sweep::canonize_to_first(mps);
    std::vector<Boundary> left(L + 1), right(L + 1);
    left[0] = mps.left_boundary();
    right[L] = mps.right_boundary();
    for (int i = L - 1; i >= 0; --i) {
        right[i] = eng.overlap_mpo_right_step(mps[i], mps[i], right[i + 1], mpo[i]);
        if (prm.spill && i + 1 < L) eng.evict(right[i + 1]);
    }
    const int northo = prm.ortho ? (int)prm.ortho->states.size() : 0;
    std::vector<std::vector<block_matrix>> oleft((size_t)northo, std::vector<block_matrix>((size_t)L + 1)), oright = oleft;
    for (int n = 0; n < northo; ++n) {
        MPS const& om = prm.ortho->states[(size_t)n];
        if ((int)om.size() != L) throw std::runtime_error("ts_sweeps: orthogonal state of a different length");
        oleft[(size_t)n][0] = mps.left_boundary()[0];
        oright[(size_t)n][(size_t)L] = mps.right_boundary()[0];
        for (int i = L - 1; i >= 0; --i) oright[(size_t)n][(size_t)i] = overlap_right_step(eng, prm.ortho->su2, mps[i], om[(size_t)i], oright[(size_t)n][(size_t)i + 1]);
    }
    if (init_seconds) *init_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init).count();
    auto to_site = [L](int i) { return i < L - 1 ? i : 2 * L - 2 - i; };      // ts_optimize.hpp:47-52: the last bond is visited twice, the first once
    bool stopped = false;
    for (int sw = 0; sw < nsweeps && !stopped; ++sw) {
        auto t0 = std::chrono::steady_clock::now();
        for (int _site = 0; _site < 2 * L - 2; ++_site) {
            if (prm.max_micro_iterations > 0 && sw == nsweeps - 1 && _site >= prm.max_micro_iterations) break;
            if (prm.should_stop && prm.should_stop()) { stopped = true; break; }
            int lr, site1, site2;
            if (_site < L - 1) { lr = 1; site1 = to_site(_site); site2 = site1 + 1; }
            else { lr = -1; site2 = to_site(_site); site1 = site2 - 1; }
            auto c0 = std::chrono::steady_clock::now();
            auto lap = [&c0]() { auto n = std::chrono::steady_clock::now(); double s = std::chrono::duration<double>(n - c0).count(); c0 = n; return s; };
            if (prm.spill) {
                eng.prefetch(left[site1]); eng.prefetch(right[site2 + 1]);
                if (lr == +1 && site2 + 2 <= L) eng.prefetch(right[site2 + 2]);
                if (lr == -1 && site1 >= 1) eng.prefetch(left[site1 - 1]);
            }
            TwoSiteTensor tst(symm, mps[site1], mps[site2]);
            MPSTensor twin = tst.make_mps();
            log.phase_seconds[0] += lap();
            MPOTensor const& tsw = ts_mpo(site1);
            log.phase_seconds[1] += lap();
            std::vector<MPSTensor> ortho_vecs((size_t)northo);
            for (int n = 0; n < northo; ++n) {
                MPS const& om = prm.ortho->states[(size_t)n];
                TwoSiteTensor ts_ortho(symm, om[(size_t)site1], om[(size_t)site2]);
                ortho_vecs[(size_t)n] = sweep::site_ortho_boundaries(twin, ts_ortho.make_mps(), oleft[(size_t)n][(size_t)site1], oright[(size_t)n][(size_t)site2 + 1]);
            }
            sweep::JDResult r = sweep::jacobi_davidson(eng, twin, left[site1], right[site2 + 1], tsw, prm.jcd_maxiter, prm.jcd_tol, ortho_vecs);
            log.phase_seconds[2] += lap();
            tst << r.vec;
            log.energies.push_back(r.theta + mpo.core_energy);
            log.n_sigma.push_back(r.n_sigma); log.total_sigma += r.n_sigma;
            Truncation trunc;
            SplitClock sclk;
            // all ranks of a sharded run must hold the same state here: same solver history, same structure about to be split
            {
                uint64_t fp = 1469598103934665603ull;
                auto mix = [&](uint64_t x) { fp ^= x; fp *= 1099511628211ull; };
                mix((uint64_t)r.n_sigma); mix((uint64_t)_site);
                r.vec.make_left_paired();
                for (auto const& q : r.vec.data().basis()) { mix(q.ls); mix(q.rs); }
                double th = r.theta; uint64_t bits; std::memcpy(&bits, &th, 8); mix(bits);
                eng.assert_consistent(fp, "eigensolver result");
            }
            if (lr == +1) {
                if (prm.alpha != 0.) tst.predict_split_l2r(prm.Mmax, prm.cutoff, prm.alpha, left[site1], mpo[site1], eng, mps[site1], mps[site2], trunc);
                else tst.split_mps_l2r(prm.Mmax, prm.cutoff, mps[site1], mps[site2], trunc, &eng);
                { double* ss = split_seconds(); sclk.lap(3); ss[3] -= 0; }
                block_matrix t = sweep::normalize_left(mps[site2]);
                if (site2 < L - 1) sweep::multiply_from_left(mps[site2 + 1], t);
                sclk.lap(4);
                log.phase_seconds[3] += lap();
                // the stale boundary at this bond goes first: the new one has the same size and takes over its memory
                if (prm.drop_stale && site2 < L - 1) right[site2] = Boundary();
                left[site2] = eng.overlap_mpo_left_step(mps[site1], mps[site1], left[site1], mpo[site1]);
                for (int n = 0; n < northo; ++n)
                    oleft[(size_t)n][(size_t)site2] = overlap_left_step(eng, prm.ortho->su2, mps[site1], prm.ortho->states[(size_t)n][(size_t)site1], oleft[(size_t)n][(size_t)site1]);
                if (prm.spill && site1 > 0) eng.evict(left[site1]);
            } else {
                if (prm.alpha != 0.) tst.predict_split_r2l(prm.Mmax, prm.cutoff, prm.alpha, right[site2 + 1], mpo[site2], eng, mps[site1], mps[site2], trunc);
                else tst.split_mps_r2l(prm.Mmax, prm.cutoff, mps[site1], mps[site2], trunc, &eng);
                sclk.lap(3);
                block_matrix t = sweep::normalize_right(mps[site1]);
                if (site1 > 0) sweep::multiply_from_right(mps[site1 - 1], t);
                sclk.lap(4);
                log.phase_seconds[3] += lap();
                if (prm.drop_stale && site1 > 0) left[site2] = Boundary();
                right[site2] = eng.overlap_mpo_right_step(mps[site2], mps[site2], right[site2 + 1], mpo[site2]);
                for (int n = 0; n < northo; ++n)
                    oright[(size_t)n][(size_t)site2] = overlap_right_step(eng, prm.ortho->su2, mps[site2], prm.ortho->states[(size_t)n][(size_t)site2], oright[(size_t)n][(size_t)site2 + 1]);
                if (prm.spill && site2 + 1 < L) eng.evict(right[site2 + 1]);
            }
            log.phase_seconds[4] += lap();
            {
                uint64_t fp = 1469598103934665603ull;
                auto mix = [&](uint64_t x) { fp ^= x; fp *= 1099511628211ull; };
                mix(trunc.bond_dimension);
                for (auto const& e : mps[site1].col_dim()) mix(e.second);
                eng.assert_consistent(fp, "bond structure after the split");
            }
            if (bond_dims) bond_dims->push_back(trunc.bond_dimension);
        }
        log.sweep_energy.push_back(log.energies.empty() ? 0. : log.energies.back());
        log.sweep_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    return log;
}

}} // namespace qcm::ts
