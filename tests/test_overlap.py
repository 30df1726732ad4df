"""MPS-MPS overlap steps (SURVEY 8(a9) remainder): contraction::Engine::overlap_left_step / overlap_right_step
(contractions/common/move_boundary.hpp:21-64), the operator-free boundary propagation behind the overlaps with orthogonal
states (optimize/optimize.h:105-117, SweepBasedAlgorithms/OverlapPropagator.h:116-124) and norm / overlap of two MPS
(mp_tensors/mps_mpo_ops.h:140-180).  The product runs them as the boundary step of the hot path with the one-entry identity
MPO tensor (qcm/overlap.hpp); the oracle restates the reference's gemm - reshape - gemm literally."""
import ctypes
import pytest
from conftest import golden


def _overlap_parity(h, symm, engine, Mbra=10, Mket=14, seed=5):
    out = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
    rc = h.lib.qcmt_overlap_parity(golden("synth_6o6e.fcidump"), symm.encode(), 6, 6, Mbra, Mket, seed, engine, out, err, 1024)
    assert rc == 0, err.value.decode()
    return list(out)


def _check(out):
    n, st, diff, left_chain, right_chain, oracle, norm2 = out[:7]
    assert n == 12                                # 6 left steps + 6 right steps, bra and ket with different bond dimensions
    assert st == 1, "block structure differs from the oracle's"
    assert diff < 1e-12, diff
    assert abs(left_chain) > 1e-8                 # a non-trivial overlap of two random states
    assert abs(left_chain - oracle) < 1e-14 + 1e-10 * abs(oracle)
    assert abs(left_chain - right_chain) < 1e-14 + 1e-10 * abs(oracle)   # <bra|ket> from either end of the chain
    assert abs(norm2 - 1.0) < 1e-12               # norm of a canonised state (mps_mpo_ops.h:140-148)


@pytest.mark.parametrize("symm", ["2u1", "su2u1", "2u1pg", "su2u1pg"])
def test_overlap_steps_plan_interpreter_vs_oracle(harness_cpu, symm):
    _check(_overlap_parity(harness_cpu, symm, 0))


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["2u1", "su2u1", "2u1pg", "su2u1pg"])
def test_overlap_steps_gpu_vs_oracle(harness_gpu, symm):
    _check(_overlap_parity(harness_gpu, symm, 1))
