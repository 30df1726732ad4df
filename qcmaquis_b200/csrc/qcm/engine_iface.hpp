// Abstract form of the reference's all-static contraction::Engine<Matrix, OtherMatrix, SymmGroup>
// (dmrg/mp_tensors/contractions/abelian/engine.hpp:26-230, non-abelian/engine.hpp:27-227) for the three
// calls on the hot path. Host code (sweep drivers, tests, benches) is written against this interface; the
// B200 implementation is qcm::GpuEngine (engine_gpu.hpp), the CPU checker lives under oracle/.
#pragma once
#include "mpo.hpp"
#include "mps.hpp"

namespace qcm {

// result of the site eigensolver (ietl::jacobi_davidson::calculate_eigenvalue, ietl/jacobi.h:361-451)
struct EigenResult { double theta = 0; MPSTensor vec; int n_sigma = 0; double resid = 0; };

struct EngineIface
{
    virtual ~EngineIface() {}
    // Optional: the whole Jacobi-Davidson solve of one site problem inside the engine (solver vectors resident on the
    // device).  false: not provided for this problem, the caller runs the host solver on site_hamil2.
    virtual bool jacobi_davidson(MPSTensor const& /*x0*/, Boundary const& /*left*/, Boundary const& /*right*/, MPOTensor const& /*mpo*/,
                                 int /*max_iter*/, double /*tol*/, EigenResult& /*res*/) { return false; }
    // engine.hpp:196-209 -- returns a LEFT-paired tensor with phys_i/left_i/right_i of the bra (== ket here)
    virtual MPSTensor site_hamil2(MPSTensor ket_tensor, Boundary const& left, Boundary const& right,
                                  MPOTensor const& mpo, bool isHermitian = true) = 0;
    // engine.hpp:102-122
    virtual Boundary overlap_mpo_left_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& left,
                                           MPOTensor const& mpo, bool isHermitian = true) = 0;
    virtual Boundary overlap_mpo_right_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& right,
                                            MPOTensor const& mpo, bool isHermitian = true) = 0;
    // ---- noise term of the perturbed density matrix (the "remaining" contractions of SURVEY 8(a9)):
    //   noise_left:  sum over b2 of Y[b2] Y[b2]^T,  Y = Engine::left_boundary_tensor_mpo(mps, left, mpo)   (move_boundary.hpp:68-95)
    //   noise_right: sum over b1 of Y'[b1]^T Y'[b1], Y' = Engine::right_boundary_tensor_mpo(mps, right, mpo) (move_boundary.hpp:97-126)
    // i.e. what prediction.hpp:34-47,101-114 and twositetensor.hpp:192-219,260-287 add, times alpha, to the reduced density matrix
    // before heev_truncate.  All blocks are returned; the caller keeps those its density matrix has.
    virtual block_matrix noise_left(MPSTensor const& /*mps*/, Boundary const& /*left*/, MPOTensor const& /*mpo*/) { throw std::runtime_error("this engine does not provide the noise term"); }
    virtual block_matrix noise_right(MPSTensor const& /*mps*/, Boundary const& /*right*/, MPOTensor const& /*mpo*/) { throw std::runtime_error("this engine does not provide the noise term"); }
    // ---- boundary storage protocol of the sweep drivers (utils/storage.h:113-185: storage::disk::prefetch / evict; drop is the
    // destruction of the Boundary).  Engines that keep boundaries in a fast tier move them here; the default does nothing.
    // make the dense blocks of a boundary this engine returned readable on the host (engines that keep results in a fast tier)
    virtual void fetch(Boundary& /*b*/) {}
    virtual void prefetch(Boundary const& /*b*/) {}
    virtual void evict(Boundary const& /*b*/) {}
    // ---- one process per GPU: the host side of a sweep runs on every rank.  Work that is the same on all ranks (the block
    // SVDs of the two-site split) is divided among them and the pieces are combined with allreduce_sum (every rank adds its
    // pieces into a zero buffer, so all ranks end up with bit-identical data); assert_consistent makes a sweep fail loudly,
    // on every rank, when the ranks' host states have diverged (a collective with mismatched sizes would hang instead).
    virtual int comm_rank() const { return 0; }
    virtual int comm_world() const { return 1; }
    virtual void allreduce_sum(double* /*buf*/, size_t /*n*/) {}
    virtual void assert_consistent(uint64_t /*fingerprint*/, const char* /*what*/) {}
};

} // namespace qcm
