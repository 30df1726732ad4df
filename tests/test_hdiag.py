"""Engine::diagonal_hamiltonian (abelian/h_diag.hpp:41-168, non-abelian/h_diag.hpp:19-155): the plan
(V = sum over b1 on the W kernels, one dense product V * DR per block) against the oracle's literal restatement, on every
site (single-site tensors) and every bond (two-site tensors) of a random MPS.  Block structure identical, values 1e-10."""
import pytest

INTERP, GPU = 0, 1
CASES = [("lih_4o.fcidump", 4, 2, 20), ("synth_6o6e.fcidump", 6, 6, 20), ("benzene_6o.fcidump", 6, 6, 30)]


@pytest.mark.parametrize("f,L,ne,M", CASES)
@pytest.mark.parametrize("symm", ["su2u1", "2u1", "su2u1pg", "2u1pg"])
def test_hdiag_plan_matches_the_oracle(harness_cpu, f, L, ne, M, symm):
    out = harness_cpu.hdiag_parity(f, symm, L, ne, M, INTERP)
    assert out[0] == 2 * L - 1 and out[1] == 1 and out[2] < 1e-12 and out[3] > 0, out[:4]


@pytest.mark.gpu
@pytest.mark.parametrize("f,L,ne,M", CASES + [("synth_6o6e.fcidump", 6, 6, 3)])
@pytest.mark.parametrize("symm", ["su2u1", "2u1", "su2u1pg"])
def test_hdiag_gpu_matches_the_oracle(harness_gpu, f, L, ne, M, symm):
    out = harness_gpu.hdiag_parity(f, symm, L, ne, M, GPU)
    assert out[0] == 2 * L - 1 and out[1] == 1 and out[2] < 1e-10 and out[3] > 0, out[:4]
