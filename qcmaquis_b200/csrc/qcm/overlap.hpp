// MPS-MPS overlap steps of the reference's contraction::Engine -- overlap_left_step / overlap_right_step
// (contractions/common/move_boundary.hpp:21-64; abelian/engine.hpp:65-82, non-abelian/engine.hpp:65-83): the boundary
// propagation WITHOUT an operator, used for the overlaps with orthogonal states (optimize.h:105-117, OverlapPropagator.h:116-124)
// and for norm / overlap of two MPS (mps_mpo_ops.h:140-180).
//
// Here they are the boundary step of the hot path with the one-entry identity MPO tensor: the same plan, the same kernels.
// A reference "overlap" block_matrix is (bra index) x (ket index); the Boundary entries of the MPO steps are stored
// transposed (ket x bra, move_boundary.hpp:65-66), so the operand goes in and the result comes out transposed.
#pragma once
#include "engine_iface.hpp"

namespace qcm {

// W = 1 x 1 tensor holding the identity on phys_i (spin-0 operator, no Hermitian partner)
inline MPOTensor identity_mpo_tensor(Index const& phys_i, bool su2)
{
    SiteOperator id;
    for (size_t s = 0; s < phys_i.size(); ++s) {
        Matrix m(phys_i[s].second, phys_i[s].second, 0.);
        for (size_t i = 0; i < m.rows; ++i) m(i, i) = 1.;
        id.bm.insert_block(m, phys_i[s].first, phys_i[s].first);
    }
    std::shared_ptr<OPTable> tbl(new OPTable());
    const tag_type tag = tbl->register_op(id);
    std::vector<PreTerm> terms(1, PreTerm{0, 0, tag, 1.0});
    std::vector<SpinDescriptor> spins(1, SpinDescriptor(0, 0, 0));
    return MPOTensor(1, 1, terms, tbl, Hermitian(1, 1), spins, spins, su2);
}

// overlap_left_step(bra, ket, left): out(bra right index, ket right index) = bra^T (left (x) 1) ket
inline block_matrix overlap_left_step(EngineIface& eng, bool su2, MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, block_matrix const& left)
{
    if (!(bra_tensor.site_dim() == ket_tensor.site_dim())) throw std::runtime_error("overlap_left_step: bra and ket differ in their physical index");
    MPOTensor id = identity_mpo_tensor(ket_tensor.site_dim(), su2);
    Boundary in; in.resize(1); in[0] = transposed(left);
    Boundary out = eng.overlap_mpo_left_step(bra_tensor, ket_tensor, in, id, false);
    eng.fetch(out);
    return transposed(out[0]);
}
// overlap_right_step(bra, ket, right): out(bra left index, ket left index)
inline block_matrix overlap_right_step(EngineIface& eng, bool su2, MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, block_matrix const& right)
{
    if (!(bra_tensor.site_dim() == ket_tensor.site_dim())) throw std::runtime_error("overlap_right_step: bra and ket differ in their physical index");
    MPOTensor id = identity_mpo_tensor(ket_tensor.site_dim(), su2);
    Boundary in; in.resize(1); in[0] = transposed(right);
    Boundary out = eng.overlap_mpo_right_step(bra_tensor, ket_tensor, in, id, false);
    eng.fetch(out);
    return transposed(out[0]);
}

// <bra | ket> through the chain of left steps (mps_mpo_ops.h:150-165 overlap); norm(mps) = overlap(mps, mps) (:140-148)
inline double overlap(EngineIface& eng, bool su2, MPS const& bra, MPS const& ket)
{
    if (bra.size() != ket.size()) throw std::runtime_error("overlap: chains of different length");
    block_matrix left = bra.left_boundary()[0];
    for (size_t i = 0; i < ket.size(); ++i) left = overlap_left_step(eng, su2, bra[i], ket[i], left);
    return left.trace();
}

} // namespace qcm
