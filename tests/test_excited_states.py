"""Excited states by orthogonal-state projection -- the callers of Engine::overlap_left_step / overlap_right_step on the sweep
path: optimize/optimize.h:44-75,105-117 (ortho_mps and their overlap boundaries, moved with every site), ss_optimize.hpp:107-111
(contraction::site_ortho_boundaries, contractions/abelian/special.hpp:16-46), ietl_lanczos_solver.h:90-95 (SingleSiteVS::project) and
the three projection points of the Jacobi-Davidson driver (ietl/jacobi.h:378,393,432).  State 0 is optimised first, state k
orthogonal to states 0 .. k-1; the three lowest energies must equal the numpy full-CI spectrum (tests/fci_numpy.py -- nothing shared
with the C++ side): the three lowest states of the Sz = 0 sector for 2u1, the three lowest singlets for su2u1."""
import ctypes, os
import pytest
from conftest import golden, GOLDEN
from fci_numpy import fci_ground_state_energy

ORACLE, INTERP, GPU = -1, 0, 1
CASES = [("synth_4o4e.fcidump", 4, 4, 16), ("synth_6o6e.fcidump", 6, 6, 64)]


def _excited(h, f, symm, L, ne, M, engine, twosite, nstates=3):
    """twosite: two-site sweeps from M0 = 4 random states (ts_optimize.hpp:120-128: the orthogonal states enter as two-site tensors,
    SU2: spin-coupled by make_mps); otherwise single-site sweeps with the noise-perturbed subspace expansion"""
    e = (ctypes.c_double * 8)(); o = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
    rc = h.lib.qcmt_excited_states_driver(golden(f), symm.encode(), L, ne, M, 6 if twosite else 8, nstates, engine, int(twosite), e, o, err, 1024)
    assert rc == 0, err.value.decode()
    return list(e[:nstates]), list(o[:nstates])


def _check(h, f, L, ne, M, symm, engine, twosite=False):
    ref = fci_ground_state_energy(os.path.join(GOLDEN, f), total_spin=0 if symm.startswith("su2") else None, n_states=3)
    e, ov = _excited(h, f, symm, L, ne, M, engine, twosite)
    assert ref[1] - ref[0] > 1e-3 and ref[2] - ref[1] > 1e-3          # three distinct levels
    for k in range(3):
        assert abs(e[k] - ref[k]) < 1e-8, (k, e, ref)
    assert max(ov[1:]) < 1e-8                                          # excited states are orthogonal to the ground state


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("f,L,ne,M", CASES)
def test_three_lowest_states_equal_the_fci_spectrum_oracle(harness_cpu, f, L, ne, M, symm):
    _check(harness_cpu, f, L, ne, M, symm, ORACLE)


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("f,L,ne,M", CASES)
def test_three_lowest_states_equal_the_fci_spectrum_plan_interpreter(harness_cpu, f, L, ne, M, symm):
    _check(harness_cpu, f, L, ne, M, symm, INTERP)


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("f,L,ne,M", CASES)
def test_three_lowest_states_equal_the_fci_spectrum_gpu(harness_gpu, f, L, ne, M, symm):
    _check(harness_gpu, f, L, ne, M, symm, GPU)


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("engine", [ORACLE, INTERP])
@pytest.mark.parametrize("f,L,ne,M", CASES)
def test_three_lowest_states_two_site_sweeps(harness_cpu, f, L, ne, M, symm, engine):
    _check(harness_cpu, f, L, ne, M, symm, engine, twosite=True)


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("f,L,ne,M", CASES)
def test_three_lowest_states_two_site_sweeps_gpu(harness_gpu, f, L, ne, M, symm):
    _check(harness_gpu, f, L, ne, M, symm, GPU, twosite=True)
