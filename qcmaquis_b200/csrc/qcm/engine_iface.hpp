// Abstract form of the reference's all-static contraction::Engine<Matrix, OtherMatrix, SymmGroup>
// (dmrg/mp_tensors/contractions/abelian/engine.hpp:26-230, non-abelian/engine.hpp:27-227) for the three
// calls on the hot path. Host code (sweep drivers, tests, benches) is written against this interface; the
// B200 implementation is qcm::GpuEngine (engine_gpu.hpp), the CPU checker lives under oracle/.
#pragma once
#include "mpo.hpp"
#include "mps.hpp"

namespace qcm {

struct EngineIface
{
    virtual ~EngineIface() {}
    // engine.hpp:196-209 -- returns a LEFT-paired tensor with phys_i/left_i/right_i of the bra (== ket here)
    virtual MPSTensor site_hamil2(MPSTensor ket_tensor, Boundary const& left, Boundary const& right,
                                  MPOTensor const& mpo, bool isHermitian = true) = 0;
    // engine.hpp:102-122
    virtual Boundary overlap_mpo_left_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& left,
                                           MPOTensor const& mpo, bool isHermitian = true) = 0;
    virtual Boundary overlap_mpo_right_step(MPSTensor const& bra_tensor, MPSTensor const& ket_tensor, Boundary const& right,
                                            MPOTensor const& mpo, bool isHermitian = true) = 0;
};

} // namespace qcm
