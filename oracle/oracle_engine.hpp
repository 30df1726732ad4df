// ORACLE -- TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's hot path; never linked into the product.
//
// Restates, function by function, the reference's CPU algorithm for
//   contraction::Engine::site_hamil2            (abelian/site_hamil.hpp:23-90, non-abelian/site_hamil.hpp:28-204)
//   contraction::Engine::overlap_mpo_left_step  (common/move_boundary.hpp:128-187)
//   contraction::Engine::overlap_mpo_right_step (common/move_boundary.hpp:189-229)
// including the same threading model (OpenMP schedule(dynamic,1) over the MPO bond index with a critical
// section reduction, dmrg/utils/parallel/loops.hpp:11-26) and the same BLAS call pattern (one dgemm per matched
// block pair; axpy panels for the W application). All paths under dmrg/framework/dmrg/mp_tensors/contractions/.
//
// Parity status: PINNED through the reference's own known answers -- Wigner 9j table (tests/test_wigner.cpp),
// E(H2) and E(LiH) ground-state energies, MPO bond dimensions / Hermitian pair counts of the shipped example
// outputs, and the sigma/boundary energy identity of test_siteproblem.cpp. See tests/test_oracle_goldens.py.
// The upstream code itself cannot be built in this image (no Boost/GSL/HDF5 headers), see DESIGN.md.
//
// Containers (Index, block_matrix, MPSTensor, MPOTensor, Boundary) are the product's host data model
// (qcmaquis_b200/csrc/qcm); only the algorithm is restated here.  The sweep drivers (qcm/sweep.hpp, qcm/twosite.hpp) are
// engine agnostic host code and run unchanged on this engine: that is how the oracle reaches the reference's pinned
// end-to-end energies (tests/test_sweeps.py) and how sweeps on the GPU engine are checked per micro-iteration.
// Also here: hdiag::diagonal_hamiltonian, the literal restatement of abelian/h_diag.hpp and non-abelian/h_diag.hpp.
#pragma once
#include "qcm/engine_iface.hpp"
#include <omp.h>

namespace oracle {
using namespace qcm;

// ---------------------------------------------------------------------------------------------------------
// block GEMMs  (block_matrix_algorithms.h:48-162; non-abelian/gemm.hpp:48-204)

// C = A*B over matched sectors; optional SU2 triangle filter on (spin(lc_A), spin, spin(rc_B))
template <class BA, class BB>
void gemm(BA const& A, BB const& B, block_matrix& C, int spin_filter = -1)
{
    C.clear();
    for (size_t k = 0; k < A.n_blocks(); ++k) {
        Charge ar = A.basis().right_charge(k);
        for (auto it = B.basis().left_lower_bound(ar); it != B.basis().end() && it->lc == ar; ++it) {
            size_t mb = it - B.basis().begin();
            if (spin_filter != -1 && !su2::triangle(spin(A.basis().left_charge(k)), spin_filter, spin(it->rc))) continue;
            MatRef a = A.block(k), b = B.block(mb);
            size_t cb = C.find_block(A.basis().left_charge(k), it->rc);
            if (cb == C.n_blocks()) cb = C.insert_block(Matrix(a.rows, it->rs), A.basis().left_charge(k), it->rc);
            if (C[cb].rows != a.rows || C[cb].cols != b.cols) {   // abelian match_and_add_block growth
                Matrix tmp(a.rows, b.cols);
                dgemm(a, b, 1.0, 0.0, tmp.data(), tmp.rows);
                C.match_and_add_block(tmp, A.basis().left_charge(k), it->rc);
            } else
                dgemm(a, b, 1.0, 1.0, C[cb].data(), C[cb].rows);
        }
    }
}

template <class BA, class BB>
void gemm_trim_left(bool is_su2, BA const& A, BB const& B, block_matrix& C, Index const& ref_left_basis, std::vector<double> scales)
{
    if (scales.size() != A.n_blocks()) scales.assign(A.n_blocks(), 1.);
    C.clear();
    if (!is_su2) {   // block_matrix_algorithms.h:104-127 (first matching B block; scales unused, all 1)
        Index B_left_basis = B.left_basis();
        for (size_t k = 0; k < A.n_blocks(); ++k) {
            size_t mb = B_left_basis.position(A.basis().right_charge(k));
            if (mb == B.n_blocks()) continue;
            if (!ref_left_basis.has(A.basis().left_charge(k))) continue;
            MatRef a = A.block(k), b = B.block(mb);
            size_t nb = C.insert_block(Matrix(a.rows, b.cols), A.basis().left_charge(k), B.basis().right_charge(mb));
            dgemm(a, b, 1.0, 0.0, C[nb].data(), C[nb].rows);
        }
        return;
    }
    for (size_t k = 0; k < A.n_blocks(); ++k) {   // non-abelian/gemm.hpp:77-113
        if (!ref_left_basis.has(A.basis().left_charge(k))) continue;
        Charge ar = A.basis().right_charge(k);
        for (auto it = B.basis().left_lower_bound(ar); it != B.basis().end() && it->lc == ar; ++it) {
            size_t mb = it - B.basis().begin();
            MatRef a = A.block(k), b = B.block(mb);
            size_t cb = C.find_block(A.basis().left_charge(k), it->rc);
            if (cb == C.n_blocks()) cb = C.insert_block(Matrix(a.rows, it->rs), A.basis().left_charge(k), it->rc);
            dgemm(a, b, scales[k], 1.0, C[cb].data(), C[cb].rows);
        }
    }
}

template <class BA, class BB>
void gemm_trim_right(bool is_su2, BA const& A, BB const& B, block_matrix& C, Index const& ref_right_basis, std::vector<double> scales)
{
    if (scales.size() != B.n_blocks()) scales.assign(B.n_blocks(), 1.);
    C.clear();
    if (!is_su2) {   // block_matrix_algorithms.h:142-162
        Index A_right_basis = A.right_basis();
        for (size_t k = 0; k < B.n_blocks(); ++k) {
            size_t mb = A_right_basis.position(B.basis().left_charge(k));
            if (mb == A.n_blocks()) continue;
            if (!ref_right_basis.has(B.basis().right_charge(k))) continue;
            MatRef a = A.block(mb), b = B.block(k);
            size_t nb = C.insert_block(Matrix(a.rows, b.cols), A.basis().left_charge(mb), B.basis().right_charge(k));
            dgemm(a, b, 1.0, 0.0, C[nb].data(), C[nb].rows);
        }
        return;
    }
    for (size_t k = 0; k < A.n_blocks(); ++k) {   // non-abelian/gemm.hpp:142-177
        Charge ar = A.basis().right_charge(k);
        for (auto it = B.basis().left_lower_bound(ar); it != B.basis().end() && it->lc == ar; ++it) {
            size_t mb = it - B.basis().begin();
            if (!ref_right_basis.has(it->rc)) continue;
            MatRef a = A.block(k), b = B.block(mb);
            size_t cb = C.find_block(A.basis().left_charge(k), it->rc);
            if (cb == C.n_blocks()) cb = C.insert_block(Matrix(a.rows, it->rs), A.basis().left_charge(k), it->rc);
            dgemm(a, b, scales[mb], 1.0, C[cb].data(), C[cb].rows);
        }
    }
}

// non-abelian/gemm.hpp:179-204: closes Y[b2] with R[b2]; only the diagonal-charge product blocks survive
template <class BA, class BB>
void gemm_trim_su2(BA const& A, BB const& B, block_matrix& C, std::vector<double> const& scales, bool conjugate_a)
{
    C.clear();
    for (size_t k = 0; k < A.n_blocks(); ++k) {
        Charge al = A.basis().left_charge(k), ar = A.basis().right_charge(k);
        size_t mb = B.basis().position(ar, al);
        if (mb == B.n_blocks()) continue;
        MatRef a = A.block(k), b = B.block(mb);
        size_t cb = C.find_block(al, al);
        if (cb == C.n_blocks()) cb = C.insert_block(Matrix(a.rows, b.cols), al, al);
        dgemm(a, b, scales[conjugate_a ? k : mb], 1.0, C[cb].data(), C[cb].rows);
    }
}

// common/boundary_times_mps.hpp:19-52
template <class BM>
std::vector<double> conjugate_phases(bool is_su2, BM const& bm, MPOTensor const& mpo, size_t k, bool left, bool forward)
{
    if (!is_su2) return std::vector<double>(bm.n_blocks(), 1.);
    int S = left ? mpo.left_spin(k).get() : mpo.right_spin(k).get();
    std::vector<double> ret(bm.n_blocks());
    for (size_t b = 0; b < bm.n_blocks(); ++b) {
        double scale = su2::conjugate_correction(spin(bm.basis().left_charge(b)), spin(bm.basis().right_charge(b)), S);
        if (forward) scale *= left ? mpo.herm_info.left_phase(mpo.herm_info.left_conj(k)) : mpo.herm_info.right_phase(mpo.herm_info.right_conj(k));
        else scale *= left ? mpo.herm_info.left_phase(k) : mpo.herm_info.right_phase(k);
        ret[b] = scale;
    }
    return ret;
}

// ---------------------------------------------------------------------------------------------------------
// step 1 containers (common/boundary_times_mps.hpp:84-218, 226-385)
class BoundaryMPSProduct
{
public:
    BoundaryMPSProduct(bool su2, MPSTensor const& mps, Boundary const& left_, MPOTensor const& mpo_, Index const& ref, bool isHermitian)
        : is_su2(su2), left(left_), mpo(mpo_), data_(left_.aux_dim()), ref_left_basis(ref), isHermitian_(isHermitian)
    {
        mps.make_right_paired();
        bm = &mps.data();
        int loop_max = (int)left.aux_dim();
#pragma omp parallel for schedule(dynamic, 1)
        for (int b1 = 0; b1 < loop_max; ++b1) {
            if (mpo.num_row_non_zeros(b1) == 1) continue;   // single-use rows are deferred to at()
            multiply(b1, data_[b1]);
        }
    }
    block_matrix const& operator[](size_t k) const { return data_[k]; }
    block_matrix const& at(size_t k, block_matrix& storage) const
    {
        if (mpo.num_row_non_zeros(k) == 1) { multiply(k, storage); return storage; }
        return data_[k];
    }
    // abelian/detail.hpp:30-60,117-135 (T_basis_left): basis of T[b] without computing it
    DualIndex basis_at(size_t b, DualIndex const& mps_basis, Index const& refBasisLeft) const
    {
        if (mpo.num_row_non_zeros(b) != 1) return data_[b].basis();
        block_view A = (mpo.herm_info.left_skip(b) && isHermitian_) ? conjugate(left[mpo.herm_info.left_conj(b)]) : transpose(left[b]);
        DualIndex ret;
        for (size_t k = 0; k < A.n_blocks(); ++k) {
            if (!refBasisLeft.has(A.basis().left_charge(k))) continue;   // refBasis.left_has
            Charge ar = A.basis().right_charge(k);
            for (auto it = mps_basis.left_lower_bound(ar); it != mps_basis.end() && it->lc == ar; ++it)
                if (!ret.has(A.basis().left_charge(k), it->rc))
                    ret.insert(QnBlock(A.basis().left_charge(k), it->rc, A.basis().left_size(k), it->rs));
        }
        return ret;
    }

private:
    void multiply(size_t b1, block_matrix& out) const
    {
        if (mpo.herm_info.left_skip(b1) && isHermitian_) {
            block_matrix const& src = left[mpo.herm_info.left_conj(b1)];
            std::vector<double> scales = conjugate_phases(is_su2, src, mpo, b1, true, false);
            gemm_trim_left(is_su2, conjugate(src), plain(*bm), out, ref_left_basis, scales);
        } else
            gemm_trim_left(is_su2, transpose(left[b1]), plain(*bm), out, ref_left_basis, std::vector<double>());
    }
    bool is_su2;
    Boundary const& left;
    MPOTensor const& mpo;
    std::vector<block_matrix> data_;
    block_matrix const* bm;
    Index ref_left_basis;
    bool isHermitian_;
};

class MPSBoundaryProduct
{
public:
    MPSBoundaryProduct(bool su2, MPSTensor const& mps, Boundary const& right_, MPOTensor const& mpo_, Index const& ref, bool isHermitian)
        : is_su2(su2), right(right_), mpo(mpo_), data_(right_.aux_dim()), pop_(right_.aux_dim(), 0), ref_right_basis(ref), isHermitian_(isHermitian)
    {
        mps.make_left_paired();
        bm = &mps.data();
        int loop_max = (int)right.aux_dim();
#pragma omp parallel for schedule(dynamic, 1)
        for (int b2 = 0; b2 < loop_max; ++b2) {
            if (mpo.num_col_non_zeros(b2) == 1) continue;
            multiply(b2, data_[b2]);
        }
    }
    block_matrix const& operator[](size_t k) const { return data_[k]; }
    block_matrix const& at(size_t k) const
    {
        if (mpo.num_col_non_zeros(k) == 1 && !pop_[k]) { multiply(k, data_[k]); pop_[k] = 1; }
        return data_[k];
    }
    // gemm_trim_right_pretend (non-abelian/gemm.hpp:115-140) / gemm_trim_right_basis (abelian/detail.hpp:73-97)
    DualIndex basis_at(size_t k) const
    {
        if (mpo.num_col_non_zeros(k) != 1) return data_[k].basis();
        block_view B = (mpo.herm_info.right_skip(k) && isHermitian_) ? adjoint(right[mpo.herm_info.right_conj(k)]) : plain(right[k]);
        DualIndex const& A = bm->basis();
        DualIndex ret;
        for (size_t a = 0; a < A.size(); ++a) {
            Charge ar = A.right_charge(a);
            for (auto it = B.basis().left_lower_bound(ar); it != B.basis().end() && it->lc == ar; ++it) {
                if (!ref_right_basis.has(it->rc)) continue;
                if (!ret.has(A.left_charge(a), it->rc)) ret.insert(QnBlock(A.left_charge(a), it->rc, A.left_size(a), it->rs));
            }
        }
        return ret;
    }
    void free(size_t b1) const
    {
        for (size_t b2 = 0; b2 < mpo.col_dim(); ++b2)
            if (mpo.num_col_non_zeros(b2) == 1 && mpo.has(b1, b2)) { data_[b2].clear(); break; }
    }

private:
    void multiply(size_t b2, block_matrix& out) const
    {
        if (mpo.herm_info.right_skip(b2) && isHermitian_) {
            block_view trv = adjoint(right[mpo.herm_info.right_conj(b2)]);
            std::vector<double> scales = conjugate_phases(is_su2, trv, mpo, b2, false, true);
            gemm_trim_right(is_su2, plain(*bm), trv, out, ref_right_basis, scales);
        } else
            gemm_trim_right(is_su2, plain(*bm), plain(right[b2]), out, ref_right_basis, std::vector<double>());
    }
    bool is_su2;
    Boundary const& right;
    MPOTensor const& mpo;
    mutable std::vector<block_matrix> data_;
    mutable std::vector<char> pop_;
    block_matrix const* bm;
    Index ref_right_basis;
    bool isHermitian_;
};

// ---------------------------------------------------------------------------------------------------------
// abelian step 2 (abelian/apply_op.hpp:23-278; block_matrix/detail/alps_detail.hpp:189-224)
namespace abelian {

inline Charge delta_of(DualIndex const& b) { return fuse(b.right_charge(0), -b.left_charge(0)); }

inline void lbtm_kernel(size_t b2, block_matrix& ret, BoundaryMPSProduct const& t, MPOTensor const& mpo,
                        DualIndex const& ket_basis, DualIndex const& bra_basis, Index const& right_i, Index const& out_left_i,
                        ProductBasis const& in_right_pb, ProductBasis const& out_left_pb)
{
    Index bra_left = bra_basis.left_basis();
    // allocate (:23-70)
    for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
        size_t b1 = mpo.row_of(e);
        DualIndex T_basis = t.basis_at(b1, ket_basis, bra_left);
        if (T_basis.size() == 0) continue;
        for (auto const& term : mpo.at_entry(e)) {
            SiteOperator const& W = mpo.op(term.first);
            if (W.n_blocks() == 0) continue;
            Charge total_delta = fuse(delta_of(W.basis()), -delta_of(T_basis));
            for (size_t r = 0; r < right_i.size(); ++r) {
                Charge out_r_charge = right_i[r].first;
                Charge out_l_charge = fuse(out_r_charge, total_delta);
                if (!out_left_i.has(out_l_charge)) continue;
                if (ret.find_block(out_l_charge, out_r_charge) == ret.n_blocks())
                    ret.insert_block(Matrix(out_left_i.size_of_block(out_l_charge), right_i[r].second), out_l_charge, out_r_charge);
            }
        }
    }
    // execute (:75-139)
    for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
        size_t b1 = mpo.row_of(e);
        block_matrix local;
        block_matrix const& T = t.at(b1, local);
        if (T.n_blocks() == 0) continue;
        for (auto const& term : mpo.at_entry(e)) {
            SiteOperator const& W = mpo.op(term.first);
            if (W.n_blocks() == 0) continue;
            Charge T_delta = delta_of(T.basis());
            Charge total_delta = fuse(delta_of(W.basis()), -T_delta);
            for (size_t r = 0; r < right_i.size(); ++r) {
                Charge out_r_charge = right_i[r].first;
                Charge out_l_charge = fuse(out_r_charge, total_delta);
                if (!out_left_i.has(out_l_charge)) continue;
                size_t r_size = right_i[r].second;
                size_t o = ret.find_block(out_l_charge, out_r_charge);
                for (size_t w_block = 0; w_block < W.n_blocks(); ++w_block) {
                    Charge phys_c1 = W.basis().left_charge(w_block), phys_c2 = W.basis().right_charge(w_block);
                    Charge in_r_charge = fuse(out_r_charge, -phys_c1);
                    Charge in_l_charge = fuse(in_r_charge, -T_delta);
                    size_t t_block = T.basis().position(in_l_charge, in_r_charge);
                    if (t_block == T.basis().size()) continue;
                    size_t in_right_offset = in_right_pb(phys_c1, out_r_charge);
                    size_t out_left_offset = out_left_pb(phys_c2, in_l_charge);
                    size_t s1 = W.basis().left_size(w_block), s2 = W.basis().right_size(w_block);
                    Matrix const& wblock = W[w_block]; Matrix const& iblock = T[t_block]; Matrix& oblock = ret[o];
                    size_t ldim = T.basis().left_size(t_block);
                    for (size_t rr = 0; rr < r_size; ++rr)          // lb_tensor_mpo
                        for (size_t ss1 = 0; ss1 < s1; ++ss1)
                            for (size_t ss2 = 0; ss2 < s2; ++ss2) {
                                double alfa = wblock(ss1, ss2) * term.second;
                                double const* in = &iblock(0, in_right_offset + ss1 * r_size + rr);
                                double* out = &oblock(out_left_offset + ss2 * ldim, rr);
                                for (size_t ll = 0; ll < ldim; ++ll) out[ll] += alfa * in[ll];
                            }
                }
            }
        }
    }
}

inline void rbtm_kernel(size_t b1, block_matrix& ret, MPSBoundaryProduct const& t, MPOTensor const& mpo,
                        Index const& left_i, Index const& out_right_i, ProductBasis const& in_left_pb, ProductBasis const& out_right_pb)
{
    for (size_t b2 : mpo.row(b1)) {   // allocate (:141-188)
        DualIndex T_basis = t.basis_at(b2);
        if (T_basis.size() == 0) continue;
        for (auto const& term : mpo.at(b1, b2)) {
            SiteOperator const& W = mpo.op(term.first);
            if (W.n_blocks() == 0) continue;
            Charge total_delta = fuse(delta_of(W.basis()), -delta_of(T_basis));
            for (size_t l = 0; l < left_i.size(); ++l) {
                Charge out_l_charge = left_i[l].first;
                Charge out_r_charge = fuse(out_l_charge, -total_delta);
                if (!out_right_i.has(out_r_charge)) continue;
                if (ret.find_block(out_l_charge, out_r_charge) == ret.n_blocks())
                    ret.insert_block(Matrix(left_i[l].second, out_right_i.size_of_block(out_r_charge)), out_l_charge, out_r_charge);
            }
        }
    }
    for (size_t b2 : mpo.row(b1)) {   // execute (:190-250)
        block_matrix const& T = t.at(b2);
        if (T.n_blocks() == 0) continue;
        for (auto const& term : mpo.at(b1, b2)) {
            SiteOperator const& W = mpo.op(term.first);
            if (W.n_blocks() == 0) continue;
            Charge T_delta = delta_of(T.basis());
            Charge total_delta = fuse(delta_of(W.basis()), -T_delta);
            for (size_t l = 0; l < left_i.size(); ++l) {
                Charge out_l_charge = left_i[l].first;
                Charge out_r_charge = fuse(out_l_charge, -total_delta);
                if (!out_right_i.has(out_r_charge)) continue;
                size_t l_size = left_i[l].second;
                size_t o = ret.find_block(out_l_charge, out_r_charge);
                for (size_t w_block = 0; w_block < W.n_blocks(); ++w_block) {
                    Charge phys_c1 = W.basis().left_charge(w_block), phys_c2 = W.basis().right_charge(w_block);
                    Charge in_l_charge = fuse(out_l_charge, phys_c1);
                    Charge in_r_charge = fuse(in_l_charge, T_delta);
                    size_t t_block = T.basis().position(in_l_charge, in_r_charge);
                    if (t_block == T.basis().size()) continue;
                    size_t in_left_offset = in_left_pb(phys_c1, out_l_charge);
                    size_t out_right_offset = out_right_pb(phys_c2, in_r_charge);
                    size_t s1 = W.basis().left_size(w_block), s2 = W.basis().right_size(w_block);
                    Matrix const& wblock = W[w_block]; Matrix const& iblock = T[t_block]; Matrix& oblock = ret[o];
                    size_t rdim = T.basis().right_size(t_block);
                    for (size_t ss1 = 0; ss1 < s1; ++ss1)           // rb_tensor_mpo
                        for (size_t ss2 = 0; ss2 < s2; ++ss2) {
                            double alfa = wblock(ss1, ss2) * term.second;
                            for (size_t rr = 0; rr < rdim; ++rr) {
                                double const* in = &iblock(in_left_offset + ss1 * l_size, rr);
                                double* out = &oblock(0, out_right_offset + ss2 * rdim + rr);
                                for (size_t ll = 0; ll < l_size; ++ll) out[ll] += in[ll] * alfa;
                            }
                        }
                }
            }
        }
    }
    t.free(b1);
}

} // namespace abelian

// ---------------------------------------------------------------------------------------------------------
// SU2 step 2 (non-abelian/apply_op.hpp, apply_op_rp.hpp, micro_kernels.hpp)
namespace nonabelian {

inline int casenr(SparseEntry const& e)
{
    if (e.row_spin == 2 && e.col_spin == 2) return 3;
    if (e.row_spin == 2) return 1;
    if (e.col_spin == 2) return 2;
    return 0;
}

// apply_op.hpp:23-101 + micro_kernels.hpp:19-45
inline void lbtm_kernel(size_t b2, block_matrix& ret, BoundaryMPSProduct const& t, MPOTensor const& mpo,
                        DualIndex const& ket_basis /*transposed*/, Index const& right_i, Index const& out_left_i,
                        ProductBasis const& in_right_pb, ProductBasis const& out_left_pb)
{
    for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
        size_t b1 = mpo.row_of(e);
        block_matrix local;
        block_matrix const& T = t.at(b1, local);
        for (auto const& term : mpo.at_entry(e)) {
            SiteOperator const& W = mpo.op(term.first);
            int a = mpo.left_spin(b1).get(), k = W.spin().get(), ap = mpo.right_spin(b2).get();
            for (size_t t_block = 0; t_block < T.n_blocks(); ++t_block) {
                Charge lc = T.basis().left_charge(t_block), rc = T.basis().right_charge(t_block);
                Charge mc = ket_basis.left_lower_bound(rc)->rc;
                for (size_t w_block = 0; w_block < W.basis().size(); ++w_block) {
                    Charge phys_in = W.basis().left_charge(w_block), phys_out = W.basis().right_charge(w_block);
                    Charge out_r_charge = fuse(rc, phys_in);
                    size_t rb = right_i.position(out_r_charge);
                    if (rb == right_i.size()) continue;
                    Charge out_l_charge = fuse(lc, phys_out);
                    if (!su2::triangle(spin(out_r_charge), ap, spin(out_l_charge))) continue;
                    if (!out_left_i.has(out_l_charge)) continue;
                    size_t r_size = right_i[rb].second;
                    size_t o = ret.find_block(out_l_charge, out_r_charge);
                    if (o == ret.n_blocks())
                        o = ret.insert_block(Matrix(out_left_i.size_of_block(out_l_charge), r_size), out_l_charge, out_r_charge);
                    int i = spin(lc), ip = spin(out_l_charge), j = spin(mc), jp = spin(out_r_charge);
                    int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                    double couplings[4];
                    su2::set_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip, term.second, couplings);
                    size_t in_right_offset = in_right_pb(phys_in, out_r_charge);
                    size_t out_left_offset = out_left_pb(phys_out, lc);
                    size_t l_size = T.basis().left_size(t_block);
                    Matrix const& iblock = T[t_block]; Matrix& oblock = ret[o];
                    for (size_t rr = 0; rr < r_size; ++rr)
                        for (int s = W.sparse_ptr[w_block]; s < W.sparse_ptr[w_block + 1]; ++s) {
                            SparseEntry const& en = W.sparse[s];
                            double alfa = en.coefficient * couplings[casenr(en)];
                            double const* in = &iblock(0, in_right_offset + en.row * r_size + rr);
                            double* out = &oblock(out_left_offset + en.col * l_size, rr);
                            for (size_t ll = 0; ll < l_size; ++ll) out[ll] += alfa * in[ll];
                        }
                }
            }
        }
    }
}

// apply_op_rp.hpp:20-101 + micro_kernels.hpp:48-98 (right-paired output, used for MPO columns with > 3 entries)
inline void lbtm_kernel_rp(size_t b2, block_matrix& ret, BoundaryMPSProduct const& t, MPOTensor const& mpo,
                           DualIndex const& ket_basis /*transposed*/, Index const& right_i, ProductBasis const& in_right_pb)
{
    for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
        size_t b1 = mpo.row_of(e);
        block_matrix local;
        block_matrix const& T = t.at(b1, local);
        for (auto const& term : mpo.at_entry(e)) {
            SiteOperator const& W = mpo.op(term.first);
            int a = mpo.left_spin(b1).get(), k = W.spin().get(), ap = mpo.right_spin(b2).get();
            for (size_t t_block = 0; t_block < T.n_blocks(); ++t_block) {
                Charge lc = T.basis().left_charge(t_block), rc = T.basis().right_charge(t_block);
                Charge mc = ket_basis.left_lower_bound(rc)->rc;
                for (size_t w_block = 0; w_block < W.basis().size(); ++w_block) {
                    Charge phys_in = W.basis().left_charge(w_block), phys_out = W.basis().right_charge(w_block);
                    Charge out_r_charge = fuse(rc, phys_in);
                    size_t rb = right_i.position(out_r_charge);
                    if (rb == right_i.size()) continue;
                    Charge out_l_charge = fuse(lc, phys_out);
                    if (!su2::triangle(spin(out_r_charge), ap, spin(out_l_charge))) continue;
                    if (!right_i.has(out_l_charge)) continue;
                    Charge out_r_charge_rp = fuse(out_r_charge, -phys_out);
                    size_t r_size = right_i[rb].second;
                    size_t o = ret.find_block(lc, out_r_charge_rp);
                    if (o == ret.n_blocks())
                        o = ret.insert_block(Matrix(T.basis().left_size(t_block), in_right_pb.size(fuse(-phys_out, out_r_charge))), lc, out_r_charge_rp);
                    int i = spin(lc), ip = spin(out_l_charge), j = spin(mc), jp = spin(out_r_charge);
                    int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                    double couplings[4];
                    su2::set_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip, term.second, couplings);
                    size_t in_right_offset = in_right_pb(phys_in, out_r_charge);
                    size_t out_right_offset = in_right_pb(phys_out, out_r_charge);
                    size_t l_size = T.basis().left_size(t_block);
                    Matrix const& iblock = T[t_block]; Matrix& oblock = ret[o];
                    size_t blength = r_size * l_size;
                    for (int s = W.sparse_ptr[w_block]; s < W.sparse_ptr[w_block + 1]; ++s) {   // rbtm_blocked
                        SparseEntry const& en = W.sparse[s];
                        double alfa = en.coefficient * couplings[casenr(en)];
                        double const* in = &iblock(0, in_right_offset + en.row * r_size);
                        double* out = &oblock(0, out_right_offset + en.col * r_size);
                        for (size_t x = 0; x < blength; ++x) out[x] += alfa * in[x];
                    }
                }
            }
        }
    }
}

struct micro_task { double scale; size_t b2, k, l_size, r_size, stripe, out_offset, in_offset; };
typedef std::map<std::pair<Charge, Charge>, std::vector<micro_task>> task_map;

// apply_op.hpp:114-209 (rbtm_tasks). The r_size_cache slicing of the reference only splits a panel into
// column slices that are processed back to back; the slices are emitted here in the same order.
inline void rbtm_tasks(size_t b1, MPSBoundaryProduct const& t, MPOTensor const& mpo, DualIndex const& ket_basis,
                       Index const& left_i, Index const& out_right_i, ProductBasis const& in_left_pb, ProductBasis const& out_right_pb,
                       task_map& tasks)
{
    for (size_t b2 : mpo.row(b1)) {
        DualIndex T = t.basis_at(b2);
        for (auto const& term : mpo.at(b1, b2)) {
            SiteOperator const& W = mpo.op(term.first);
            int a = mpo.left_spin(b1).get(), k = W.spin().get(), ap = mpo.right_spin(b2).get();
            for (size_t t_block = 0; t_block < T.size(); ++t_block) {
                Charge lc = T.left_charge(t_block), rc = T.right_charge(t_block);
                Charge mc = ket_basis.left_lower_bound(lc)->rc;
                for (size_t w_block = 0; w_block < W.basis().size(); ++w_block) {
                    Charge phys_in = W.basis().left_charge(w_block), phys_out = W.basis().right_charge(w_block);
                    Charge out_l_charge = fuse(lc, -phys_in);
                    size_t lb = left_i.position(out_l_charge);
                    if (lb == left_i.size()) continue;
                    Charge out_r_charge = fuse(rc, -phys_out);
                    if (!su2::triangle(spin(out_l_charge), a, spin(out_r_charge))) continue;
                    if (!out_right_i.has(out_r_charge)) continue;
                    size_t l_size = left_i[lb].second;
                    std::vector<micro_task>& otasks = tasks[std::make_pair(out_l_charge, out_r_charge)];
                    int i = spin(out_r_charge), ip = spin(rc), j = spin(out_l_charge), jp = spin(mc);
                    int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                    double couplings[4];
                    su2::set_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip, term.second, couplings);
                    size_t in_left_offset = in_left_pb(phys_in, out_l_charge);
                    size_t out_right_offset = out_right_pb(phys_out, rc);
                    size_t r_size = T.right_size(t_block);
                    micro_task tpl; tpl.l_size = l_size; tpl.stripe = T.left_size(t_block); tpl.b2 = b2; tpl.k = t_block;
                    size_t r_size_cache = 16384 / (l_size * W.basis().right_size(w_block));
                    if (r_size_cache == 0) r_size_cache = 1;
                    auto op_iterate = [&](size_t in_offset, size_t rsc, size_t oro) {
                        for (int s = W.sparse_ptr[w_block]; s < W.sparse_ptr[w_block + 1]; ++s) {
                            SparseEntry const& en = W.sparse[s];
                            micro_task task = tpl;
                            task.in_offset = in_offset + en.row * tpl.l_size;
                            task.scale = en.coefficient * couplings[casenr(en)];
                            task.r_size = rsc;
                            task.out_offset = oro + en.col * r_size;
                            otasks.push_back(task);
                        }
                    };
                    for (size_t slice = 0; slice < r_size / r_size_cache; ++slice) {
                        size_t roc = slice * r_size_cache;
                        op_iterate(in_left_offset + tpl.stripe * roc, r_size_cache, out_right_offset + roc);
                    }
                    size_t r_size_remain = r_size % r_size_cache;
                    size_t right_offset_remain = r_size - r_size_remain;
                    if (r_size_remain == 0) continue;
                    op_iterate(in_left_offset + tpl.stripe * right_offset_remain, r_size_remain, out_right_offset + right_offset_remain);
                }
            }
        }
    }
}

inline void task_axpy(micro_task const& task, double* oblock, double const* source)
{
    for (size_t rr = 0; rr < task.r_size; ++rr) {
        double const* in = source + task.stripe * rr;
        double* out = oblock + (task.out_offset + rr) * task.l_size;
        for (size_t ll = 0; ll < task.l_size; ++ll) out[ll] += task.scale * in[ll];
    }
}

// apply_op.hpp:211-232 (rbtm_axpy) + :234-253 (rbtm_kernel)
inline void rbtm_kernel(size_t b1, block_matrix& ret, MPSBoundaryProduct const& t, MPOTensor const& mpo, DualIndex const& ket_basis,
                        Index const& left_i, Index const& out_right_i, ProductBasis const& in_left_pb, ProductBasis const& out_right_pb)
{
    task_map tasks;
    rbtm_tasks(b1, t, mpo, ket_basis, left_i, out_right_i, in_left_pb, out_right_pb, tasks);
    for (auto& kv : tasks) {
        std::vector<micro_task>& otasks = kv.second;
        std::stable_sort(otasks.begin(), otasks.end(), [](micro_task const& x, micro_task const& y) { return x.out_offset < y.out_offset; });
        if (otasks.empty()) continue;
        Matrix buf(otasks[0].l_size, out_right_i.size_of_block(kv.first.second));
        for (auto const& task : otasks) task_axpy(task, buf.data(), t.at(task.b2)[task.k].data() + task.in_offset);
        ret.insert_block(buf, kv.first.first, kv.first.second);
    }
    t.free(b1);
}

// apply_op.hpp:255-298 (charge_gemm + rbtm_axpy_gemm)
inline void rbtm_axpy_gemm(size_t b1, task_map& tasks, block_matrix& prod, Index const& out_right_i, MPOTensor const& mpo,
                           block_view const& left_b1, MPSBoundaryProduct const& t)
{
    // NOTE the reference keys the phase lookup on left_skip(b1) alone (apply_op.hpp:279-280)
    std::vector<double> phases = mpo.herm_info.left_skip(b1) ? conjugate_phases(true, left_b1, mpo, b1, true, false)
                                                             : std::vector<double>(left_b1.n_blocks(), 1.);
    for (auto& kv : tasks) {
        std::vector<micro_task>& otasks = kv.second;
        std::stable_sort(otasks.begin(), otasks.end(), [](micro_task const& x, micro_task const& y) { return x.out_offset < y.out_offset; });
        if (otasks.empty()) continue;
        Matrix buf(otasks[0].l_size, out_right_i.size_of_block(kv.first.second));
        size_t k = left_b1.basis().position(kv.first.second, kv.first.first);
        if (k == left_b1.basis().size()) continue;
        for (auto const& task : otasks) task_axpy(task, buf.data(), t.at(task.b2)[task.k].data() + task.in_offset);
        Charge rc = kv.first.second;
        size_t cb = prod.find_block(rc, rc);
        MatRef a = left_b1.block(k);
        if (cb == prod.n_blocks()) cb = prod.insert_block(Matrix(a.rows, buf.cols), rc, rc);
        MatRef b{buf.data(), buf.rows, buf.cols, buf.rows, false};
        dgemm(a, b, phases[k], 1.0, prod[cb].data(), prod[cb].rows);
    }
}

} // namespace nonabelian

// ---------------------------------------------------------------------------------------------------------
// diagonal_hamiltonian: diag(H_eff) in the left-paired layout of x (preconditioner of the Davidson solver,
// optimize/ietl_davidson.h:85-114).  Restated literally, quirks included:
//   abelian  (abelian/h_diag.hpp:41-117,133-168): only op(0) of every MPO entry, the entry's scale is not applied;
//   SU2      (non-abelian/h_diag.hpp:19-113,115-155): all ops, couplings from mod_coupling with the 2x2 case table.
// Both read the STORED boundary blocks only: a Hermitian-skipped bond (empty entry) contributes nothing.
namespace hdiag {

inline block_matrix lbtm_diag_kernel(bool su2, size_t b2, Boundary const& left, MPOTensor const& mpo, Index const& out_left_i, Index const& left_i,
                                     Index const& right_i, Index const& phys_i, ProductBasis const& left_pb)
{
    block_matrix ret;
    for (size_t e = mpo.col_begin(b2); e < mpo.col_end(b2); ++e) {
        size_t b1 = mpo.row_of(e);
        auto const& access = mpo.at_entry(e);
        size_t n_ops = su2 ? access.size() : 1;
        for (size_t op_index = 0; op_index < n_ops; ++op_index) {
            SiteOperator const& W = mpo.op(access[op_index].first);
            int a = 0, k = 0, ap = 0;
            if (su2) { a = mpo.left_spin(b1).get(); k = W.spin().get(); ap = mpo.right_spin(b2).get(); }
            for (size_t block = 0; block < right_i.size(); ++block) {
                Charge in_charge = right_i[block].first;
                size_t o = ret.find_block(in_charge, in_charge);
                if (o == ret.n_blocks()) o = ret.insert_block(Matrix(out_left_i[block].second, right_i[block].second), in_charge, in_charge);
                for (size_t s = 0; s < phys_i.size(); ++s) {
                    Charge phys_charge = phys_i[s].first;
                    size_t l = left_i.position(fuse(in_charge, -phys_charge));
                    if (l == left_i.size()) continue;
                    Charge lc = left_i[l].first;
                    size_t l_block = left[b1].find_block(lc, lc);
                    if (l_block == left[b1].n_blocks()) continue;
                    Matrix const& Lb = left[b1][l_block];
                    std::vector<double> left_diagonal(left_i[l].second);
                    for (size_t i = 0; i < left_diagonal.size(); ++i) left_diagonal[i] = Lb(i, i);
                    size_t left_offset = left_pb(phys_charge, lc);
                    for (size_t w_block = 0; w_block < W.basis().size(); ++w_block) {
                        Charge phys_in = W.basis().left_charge(w_block), phys_out = W.basis().right_charge(w_block);
                        if (!(phys_charge == phys_in) || !(phys_in == phys_out)) continue;
                        double couplings[2] = {1., 1.};
                        if (su2) {
                            int i = spin(lc), ip = spin(in_charge), j = spin(lc), jp = spin(in_charge);
                            int two_sp = std::abs(i - ip), two_s = std::abs(j - jp);
                            double prefactor = std::sqrt((ip + 1.) * (j + 1.) / ((i + 1.) * (jp + 1.))) * access[op_index].second;
                            couplings[0] = prefactor * su2::mod_coupling(j, two_s, jp, a, k, ap, i, two_sp, ip);
                            couplings[1] = prefactor * su2::mod_coupling(j, 2, jp, a, k, ap, i, 2, ip);
                        }
                        for (int sp = W.sparse_ptr[w_block]; sp < W.sparse_ptr[w_block + 1]; ++sp) {
                            SparseEntry const& en = W.sparse[sp];
                            size_t ss1 = en.row;
                            if (ss1 != (size_t)en.col) continue;
                            double alfa_t = su2 ? en.coefficient * couplings[en.row_spin == 2 ? 1 : 0] : en.coefficient;
                            Matrix& out = ret[o];
                            for (size_t col_i = 0; col_i < right_i[block].second; ++col_i)
                                for (size_t i = 0; i < left_diagonal.size(); ++i)
                                    out(left_offset + ss1 * left_i[l].second + i, col_i) += left_diagonal[i] * alfa_t;
                        }
                    }
                }
            }
        }
    }
    return ret;
}

inline block_matrix diagonal_hamiltonian(SymmKind symm, Boundary const& left, Boundary const& right, MPOTensor const& mpo, MPSTensor const& x)
{
    const bool su2 = is_su2(symm);
    Index const& physical_i = x.site_dim();
    Index right_i = x.col_dim(), out_left_i = physical_i * x.row_dim();
    common_subset(out_left_i, right_i);
    ProductBasis out_left_pb(physical_i, x.row_dim());
    block_matrix ret;
    for (size_t b2 = 0; b2 < right.aux_dim(); ++b2) {
        block_matrix lb2 = lbtm_diag_kernel(su2, b2, left, mpo, out_left_i, x.row_dim(), x.col_dim(), physical_i, out_left_pb);
        for (size_t block = 0; block < lb2.n_blocks(); ++block) {
            Charge in_r_charge = lb2.basis()[block].rc;
            size_t rblock = right[b2].find_block(in_r_charge, in_r_charge);
            if (rblock == right[b2].n_blocks()) continue;
            Matrix m = lb2[block];
            Matrix const& R = right[b2][rblock];
            for (size_t c = 0; c < m.cols; ++c) for (size_t i = 0; i < m.rows; ++i) m(i, c) *= R(c, c);
            ret.match_and_add_block(m, in_r_charge, in_r_charge);
        }
    }
    return ret;
}

} // namespace hdiag

// ---------------------------------------------------------------------------------------------------------
class OracleEngine : public EngineIface
{
public:
    explicit OracleEngine(SymmKind s) : symm(s), su2_(is_su2(s)) {}

    MPSTensor site_hamil2(MPSTensor ket_tensor, Boundary const& left, Boundary const& right, MPOTensor const& mpo, bool isHermitian = true) override
    {
        if (!su2_) return site_hamil_abelian(ket_tensor, ket_tensor, left, right, mpo, isHermitian);
        // non-abelian/site_hamil.hpp:28-38: direction by the number of non-single-use bonds
        if ((mpo.row_dim() - mpo.num_one_rows()) < (mpo.col_dim() - mpo.num_one_cols()))
            return site_hamil_lbtm(ket_tensor, ket_tensor, left, right, mpo, isHermitian);
        return site_hamil_rbtm(ket_tensor, ket_tensor, left, right, mpo, isHermitian);
    }

    // exposed so that tests can check that the two SU2 variants agree
    MPSTensor site_hamil_lbtm(MPSTensor ket_tensor, MPSTensor const& bra_tensor, Boundary const& left, Boundary const& right,
                              MPOTensor const& mpo, bool isHermitian)
    {
        Index const& physical_i = ket_tensor.site_dim();
        Index const& left_i = bra_tensor.row_dim();
        Index right_i = ket_tensor.col_dim();
        Index out_left_i = physical_i * left_i;
        Index right_i_bra = bra_tensor.col_dim();
        common_subset(out_left_i, right_i_bra);
        ProductBasis out_left_pb(physical_i, left_i);
        ProductBasis in_right_pb(physical_i, right_i, true);
        bra_tensor.make_right_paired();
        Index indexForTrim = bra_tensor.data().left_basis();
        BoundaryMPSProduct t(true, ket_tensor, left, mpo, indexForTrim, isHermitian);
        MPSTensor ret;
        ret.phys_i = bra_tensor.site_dim(); ret.left_i = bra_tensor.row_dim(); ret.right_i = bra_tensor.col_dim();
        DualIndex ket_basis_transpose = swapped(ket_tensor.data().basis());
        int loop_max = (int)mpo.col_dim();
        block_matrix collector;
#pragma omp parallel for schedule(dynamic, 1)
        for (int b2 = 0; b2 < loop_max; ++b2) {
            block_matrix grid, tmp, tmp2;
            size_t num_ops = mpo.col_end(b2) - mpo.col_begin(b2);
            if (num_ops > 3) {
                nonabelian::lbtm_kernel_rp(b2, grid, t, mpo, ket_basis_transpose, right_i, in_right_pb);
                reshape_right_to_left_new(physical_i, left_i, right_i, grid, tmp2);
                std::swap(grid, tmp2);
            } else
                nonabelian::lbtm_kernel(b2, grid, t, mpo, ket_basis_transpose, right_i, out_left_i, in_right_pb, out_left_pb);
            if (mpo.herm_info.right_skip(b2) && isHermitian) {
                block_view adj = adjoint(right[mpo.herm_info.right_conj(b2)]);
                std::vector<double> phases = conjugate_phases(true, adj, mpo, b2, false, true);
                gemm_trim_su2(plain(grid), adj, tmp, phases, false);
            } else
                gemm_trim_su2(plain(grid), plain(right[b2]), tmp, std::vector<double>(grid.n_blocks(), 1.), true);
            if (num_ops > 3)
                for (size_t k = 0; k < tmp.n_blocks(); ++k)
                    if (!out_left_i.has(tmp.basis().left_charge(k))) { tmp.remove_block(k); --k; }
#pragma omp critical(oracle_sigma_reduce)
            for (size_t k = 0; k < tmp.n_blocks(); ++k)
                collector.match_and_add_block(tmp[k], tmp.basis().left_charge(k), tmp.basis().right_charge(k));
        }
        ret = MPSTensor(ret.phys_i, ret.left_i, ret.right_i, collector, LeftPaired, true);
        return ret;
    }

    MPSTensor site_hamil_rbtm(MPSTensor ket_tensor, MPSTensor const& bra_tensor, Boundary const& left, Boundary const& right,
                              MPOTensor const& mpo, bool isHermitian)
    {
        bra_tensor.make_left_paired();
        Index indexForTrim = bra_tensor.data().left_basis();
        MPSBoundaryProduct t(true, ket_tensor, right, mpo, indexForTrim, isHermitian);
        Index const& physical_i = ket_tensor.site_dim();
        Index right_i = bra_tensor.col_dim();
        Index left_i = ket_tensor.row_dim(), out_right_i = adjoin(physical_i) * right_i;
        Index left_i_ket = ket_tensor.row_dim();
        common_subset(out_right_i, left_i_ket);
        ProductBasis in_left_pb(physical_i, left_i);
        ProductBasis out_right_pb(physical_i, right_i, true);
        block_matrix collector;
        int loop_max = (int)mpo.row_dim();
        DualIndex ket_basis = ket_tensor.data().basis();
#pragma omp parallel for schedule(dynamic, 1)
        for (int b1 = 0; b1 < loop_max; ++b1) {
            block_matrix tmp2;
            nonabelian::task_map tasks;
            nonabelian::rbtm_tasks(b1, t, mpo, ket_basis, left_i, out_right_i, in_left_pb, out_right_pb, tasks);
            if (mpo.herm_info.left_skip(b1) && isHermitian)
                nonabelian::rbtm_axpy_gemm(b1, tasks, tmp2, out_right_i, mpo, conjugate(left[mpo.herm_info.left_conj(b1)]), t);
            else
                nonabelian::rbtm_axpy_gemm(b1, tasks, tmp2, out_right_i, mpo, transpose(left[b1]), t);
            t.free(b1);
#pragma omp critical(oracle_sigma_reduce)
            for (size_t k = 0; k < tmp2.n_blocks(); ++k)
                collector.match_and_add_block(tmp2[k], tmp2.basis().left_charge(k), tmp2.basis().right_charge(k));
        }
        block_matrix out;
        reshape_right_to_left_new(physical_i, left_i, right_i, collector, out);
        return MPSTensor(bra_tensor.site_dim(), bra_tensor.row_dim(), bra_tensor.col_dim(), out, LeftPaired, true);
    }

    MPSTensor site_hamil_abelian(MPSTensor ket_tensor, MPSTensor const& bra_tensor, Boundary const& left, Boundary const& right,
                                 MPOTensor const& mpo, bool isHermitian)
    {
        Index const& physical_i = ket_tensor.site_dim();
        Index const& left_i = bra_tensor.row_dim();
        bra_tensor.make_right_paired();
        Index indexForTrim = bra_tensor.data().left_basis();
        BoundaryMPSProduct t(false, ket_tensor, left, mpo, indexForTrim, isHermitian);
        Index right_i = ket_tensor.col_dim(), out_left_i = physical_i * left_i;
        common_subset(out_left_i, right_i);   // trims BOTH (abelian/site_hamil.hpp:42)
        ProductBasis out_left_pb(physical_i, left_i);
        ProductBasis in_right_pb(physical_i, right_i, true);
        DualIndex ket_basis = ket_tensor.data().basis(), bra_basis = bra_tensor.data().basis();
        block_matrix collector;
        int loop_max = (int)mpo.col_dim();
#pragma omp parallel for schedule(dynamic, 1)
        for (int b2 = 0; b2 < loop_max; ++b2) {
            block_matrix grid, tmp;
            abelian::lbtm_kernel(b2, grid, t, mpo, ket_basis, bra_basis, right_i, out_left_i, in_right_pb, out_left_pb);
            if (mpo.herm_info.right_skip(b2) && isHermitian) gemm(plain(grid), adjoint(right[mpo.herm_info.right_conj(b2)]), tmp);
            else gemm(plain(grid), plain(right[b2]), tmp);
#pragma omp critical(oracle_sigma_reduce)
            for (size_t k = 0; k < tmp.n_blocks(); ++k)
                collector.match_and_add_block(tmp[k], tmp.basis().left_charge(k), tmp.basis().right_charge(k));
        }
        return MPSTensor(bra_tensor.site_dim(), bra_tensor.row_dim(), bra_tensor.col_dim(), collector, LeftPaired, true);
    }

    // common/move_boundary.hpp:128-187
    Boundary overlap_mpo_left_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& left,
                                   MPOTensor const& mpo, bool isHermitian = true) override
    {
        bra_tensor.make_right_paired();
        Index braBasis = bra_tensor.data().left_basis();
        MPSTensor ket_cpy = ket_tensor;
        BoundaryMPSProduct t(su2_, ket_cpy, left, mpo, braBasis, isHermitian);
        Index const& left_i = bra_tensor.row_dim();
        Index right_i = ket_tensor.col_dim();
        Index bra_right_i = bra_tensor.col_dim();
        Index out_left_i = bra_tensor.site_dim() * left_i;
        common_subset(out_left_i, bra_right_i);
        ProductBasis out_left_pb(bra_tensor.site_dim(), left_i);
        ProductBasis in_right_pb(ket_tensor.site_dim(), right_i, true);
        int loop_max = (int)mpo.col_dim();
        DualIndex bra_basis = bra_tensor.data().basis();
        bra_tensor.make_left_paired();
        block_matrix bra_conj = bra_tensor.data();
        DualIndex ket_basis_transpose = swapped(ket_cpy.data().basis());
        Boundary ret; ret.resize(loop_max);
#pragma omp parallel for schedule(dynamic, 1)
        for (int b2 = 0; b2 < loop_max; ++b2) {
            if (mpo.herm_info.right_skip(b2) && isHermitian) continue;
            block_matrix grid;
            if (su2_) nonabelian::lbtm_kernel(b2, grid, t, mpo, ket_basis_transpose, right_i, out_left_i, in_right_pb, out_left_pb);
            else abelian::lbtm_kernel(b2, grid, t, mpo, ket_basis_transpose, bra_basis, right_i, out_left_i, in_right_pb, out_left_pb);
            gemm(transpose(grid), plain(bra_conj), ret[b2], su2_ ? mpo.right_spin(b2).get() : -1);
        }
        return ret;
    }

    // common/move_boundary.hpp:189-229
    Boundary overlap_mpo_right_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& right,
                                    MPOTensor const& mpo, bool isHermitian = true) override
    {
        Index const& physical_i = ket_tensor.site_dim();
        Index right_i = bra_tensor.col_dim();
        MPSTensor ket_cpy = ket_tensor;
        Index left_i = ket_tensor.row_dim(), out_right_i = adjoin(physical_i) * right_i, bra_left_i = bra_tensor.row_dim();
        bra_tensor.make_left_paired();
        Index indexForTrim = bra_tensor.data().right_basis();
        MPSBoundaryProduct t(su2_, ket_cpy, right, mpo, indexForTrim, isHermitian);
        common_subset(out_right_i, bra_left_i);
        ProductBasis in_left_pb(physical_i, left_i);
        ProductBasis out_right_pb(physical_i, right_i, true);
        Boundary ret; ret.resize(mpo.row_dim());
        int loop_max = (int)mpo.row_dim();
        bra_tensor.make_right_paired();
        block_matrix bra_conj = bra_tensor.data();
        DualIndex ket_basis = ket_cpy.data().basis();
#pragma omp parallel for schedule(dynamic, 1)
        for (int b1 = 0; b1 < loop_max; ++b1) {
            if (mpo.herm_info.left_skip(b1) && isHermitian) continue;
            block_matrix y;
            if (su2_) nonabelian::rbtm_kernel(b1, y, t, mpo, ket_basis, left_i, out_right_i, in_left_pb, out_right_pb);
            else abelian::rbtm_kernel(b1, y, t, mpo, left_i, out_right_i, in_left_pb, out_right_pb);
            gemm(plain(y), transpose(bra_conj), ret[b1], su2_ ? mpo.left_spin(b1).get() : -1);
        }
        return ret;
    }

    // common/move_boundary.hpp:21-45 (overlap_left_step): t1 = left * ket (right-paired), reshaped to left-paired with the bra's
    // row index, closed with bra^T.  Gemm::gemm is the plain block product for both symmetry kinds (SU2::gemm with spin = -1)
    block_matrix overlap_left_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, block_matrix const& left)
    {
        block_matrix t1, t3;
        ket_tensor.make_right_paired();
        gemm(plain(left), plain(ket_tensor.data()), t1);
        reshape_right_to_left_new(ket_tensor.site_dim(), bra_tensor.row_dim(), ket_tensor.col_dim(), t1, t3);
        bra_tensor.make_left_paired();
        gemm(transpose(bra_tensor.data()), plain(t3), t1);
        return t1;
    }
    // common/move_boundary.hpp:47-63 (overlap_right_step)
    block_matrix overlap_right_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, block_matrix const& right)
    {
        block_matrix t1, t3;
        ket_tensor.make_left_paired();
        gemm(plain(ket_tensor.data()), transpose(right), t1);
        reshape_left_to_right_new(ket_tensor.site_dim(), ket_tensor.row_dim(), bra_tensor.col_dim(), t1, t3);
        bra_tensor.make_right_paired();
        gemm(plain(bra_tensor.data()), transpose(t3), t1);
        return t1;
    }

    // common/move_boundary.hpp:68-95 (left_boundary_tensor_mpo) followed by the accumulation loop of prediction.hpp:34-47 /
    // twositetensor.hpp:196-219 without alpha and without the restriction to the density matrix's blocks
    block_matrix noise_left(MPSTensor const& mps_in, Boundary const& left, MPOTensor const& mpo) override
    {
        MPSTensor mps = mps_in;
        Index physical_i = mps.site_dim(), left_i = mps.row_dim(), right_i = mps.col_dim(), out_left_i = physical_i * left_i;
        BoundaryMPSProduct t(su2_, mps, left, mpo, left_i, true);
        ProductBasis out_left_pb(physical_i, left_i);
        ProductBasis in_right_pb(physical_i, right_i, true);
        mps.make_right_paired();
        DualIndex ket_basis = mps.data().basis();
        DualIndex ket_basis_transpose = swapped(ket_basis);
        const int loop_max = (int)mpo.col_dim();
        std::vector<block_matrix> parts((size_t)loop_max);
#pragma omp parallel for schedule(dynamic, 1)
        for (int b2 = 0; b2 < loop_max; ++b2) {
            block_matrix grid;
            if (su2_) nonabelian::lbtm_kernel(b2, grid, t, mpo, ket_basis_transpose, right_i, out_left_i, in_right_pb, out_left_pb);
            else abelian::lbtm_kernel(b2, grid, t, mpo, ket_basis, ket_basis, right_i, out_left_i, in_right_pb, out_left_pb);
            gemm(plain(grid), transpose(grid), parts[(size_t)b2], -1);
        }
        block_matrix ret;
        for (auto const& tdm : parts) for (size_t k = 0; k < tdm.n_blocks(); ++k) ret.match_and_add_block(tdm[k], tdm.basis().left_charge(k), tdm.basis().right_charge(k));
        return ret;
    }
    // common/move_boundary.hpp:97-126 (right_boundary_tensor_mpo) + prediction.hpp:101-114 / twositetensor.hpp:264-287
    block_matrix noise_right(MPSTensor const& mps_in, Boundary const& right, MPOTensor const& mpo) override
    {
        MPSTensor mps = mps_in;
        Index physical_i = mps.site_dim(), left_i = mps.row_dim(), right_i = mps.col_dim(), out_right_i = adjoin(physical_i) * right_i;
        MPSBoundaryProduct t(su2_, mps, right, mpo, mps.row_dim(), true);       // "constructor without index": boundary_times_mps.hpp:258-261
        ProductBasis in_left_pb(physical_i, left_i);
        ProductBasis out_right_pb(physical_i, right_i, true);
        mps.make_left_paired();
        DualIndex ket_basis = mps.data().basis();
        const int loop_max = (int)mpo.row_dim();
        std::vector<block_matrix> parts((size_t)loop_max);
#pragma omp parallel for schedule(dynamic, 1)
        for (int b1 = 0; b1 < loop_max; ++b1) {
            block_matrix y;
            if (su2_) nonabelian::rbtm_kernel(b1, y, t, mpo, ket_basis, left_i, out_right_i, in_left_pb, out_right_pb);
            else abelian::rbtm_kernel(b1, y, t, mpo, left_i, out_right_i, in_left_pb, out_right_pb);
            gemm(transpose(y), plain(y), parts[(size_t)b1], -1);
        }
        block_matrix ret;
        for (auto const& tdm : parts) for (size_t k = 0; k < tdm.n_blocks(); ++k) ret.match_and_add_block(tdm[k], tdm.basis().left_charge(k), tdm.basis().right_charge(k));
        return ret;
    }

private:
    static DualIndex swapped(DualIndex const& b)
    {
        // "ket_basis comes transposed": lc/rc swapped IN PLACE, order kept (non-abelian/site_hamil.hpp:85-89)
        DualIndex r;
        for (size_t i = 0; i < b.size(); ++i) r.push_back_unsorted(QnBlock(b[i].rc, b[i].lc, b[i].rs, b[i].ls));
        return r;
    }
    SymmKind symm;
    bool su2_;
};

} // namespace oracle
