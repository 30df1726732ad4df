"""An INDEPENDENT known answer for the systems the reference holds no printed energy for: full configuration interaction in numpy
(tests/fci_numpy.py -- no MPO, no symmetry containers, nothing shared with the product or the oracle) against DMRG sweeps that
keep the full bond dimension.  This pins term generation + MPO construction + two-site fusion + contraction + solver + split at
L = 6 and L = 8 (BASELINE config 1's system) for both symmetry families: a 2u1 run with nup = ndown lands on the lowest state of
the Sz = 0 sector, an su2u1 run with spin 0 on the lowest SINGLET (the synthetic integrals have a high-spin ground state, so the
two differ by 0.05-0.1 Eh -- a wrong spin coupling could not go unnoticed).  The FCI itself reproduces the three energies the
reference prints (test1.cpp:93, LiHFixture.h:112, H2_2e4o.TI.SS.out:70)."""
import json, os
import pytest
from conftest import GOLDEN
from fci_numpy import fci_ground_state_energy

ORACLE, INTERP, GPU = -1, 0, 1
FCI = json.load(open(os.path.join(GOLDEN, "fci_values.json")))
REF = json.load(open(os.path.join(GOLDEN, "reference_values.json")))
E_TOL = 1e-8          # north star: energies within 1e-8 Eh


def test_fci_reproduces_the_reference_held_energies():
    for name in ("h2_2o", "h2_4o", "lih_4o"):
        f = os.path.join(GOLDEN, name + ".fcidump")
        assert fci_ground_state_energy(f, total_spin=0) == pytest.approx(REF["energies"][name]["value"], abs=1e-12)


def test_stored_fci_values_are_what_the_script_computes():
    f = os.path.join(GOLDEN, "synth_6o6e.fcidump")
    assert fci_ground_state_energy(f) == pytest.approx(FCI["synth_6o6e"]["sz0_ground_state"], abs=1e-11)
    assert fci_ground_state_energy(f, total_spin=0) == pytest.approx(FCI["synth_6o6e"]["singlet_ground_state"], abs=1e-11)
    assert FCI["synth_6o6e"]["singlet_ground_state"] - FCI["synth_6o6e"]["sz0_ground_state"] > 0.01      # the spin sectors are told apart


def _key(symm):
    return "singlet_ground_state" if symm.startswith("su2") else "sz0_ground_state"


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
@pytest.mark.parametrize("engine", [ORACLE, INTERP])
def test_full_bond_dimension_sweeps_reach_the_fci_energy_6o6e(harness_cpu, symm, engine):
    e, info = harness_cpu.ts_dmrg("synth_6o6e.fcidump", symm, 6, 6, 8, 64, 3, engine)
    assert e[-1] == pytest.approx(FCI["synth_6o6e"][_key(symm)], abs=E_TOL)
    assert min(e) > FCI["synth_6o6e"][_key(symm)] - E_TOL          # variational


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_single_site_sweeps_with_noise_reach_the_fci_energy_4o4e(harness_cpu, symm):
    """single-site DMRG cannot grow its bonds without the noise term: from an M = 2 start it reaches the exact energy only
    through grow_l2r/r2l_sweep (ts::NoiseGrow)"""
    import ctypes
    from conftest import golden
    e = (ctypes.c_double * 512)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
    rc = harness_cpu.lib.qcmt_ss_dmrg_noise(golden("synth_4o4e.fcidump"), symm.encode(), 4, 4, 2, 64, 6, 42, INTERP, ctypes.c_double(1e-4), ctypes.c_double(1e-14),
                                            e, 512, ctypes.byref(n), info, err, 1024)
    assert rc == 0, err.value.decode()
    assert e[n.value - 1] == pytest.approx(FCI["synth_4o4e"][_key(symm)], abs=1e-7)
    rc = harness_cpu.lib.qcmt_ss_dmrg_noise(golden("synth_4o4e.fcidump"), symm.encode(), 4, 4, 2, 64, 6, 42, INTERP, ctypes.c_double(0.), ctypes.c_double(1e-14),
                                            e, 512, ctypes.byref(n), info, err, 1024)
    assert rc == 0, err.value.decode()
    assert e[n.value - 1] > FCI["synth_4o4e"][_key(symm)] + 1e-4        # without noise the M = 2 bonds cannot grow


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_oracle_config1_system_reaches_the_fci_energy(harness_cpu, fcidump_8o8e, symm):
    """BASELINE configs[0]'s system (8e/8o) at the full bond dimension 4^4 = 256 on the CPU oracle"""
    e, info = harness_cpu.ts_dmrg(fcidump_8o8e, symm, 8, 8, 16, 256, 6, ORACLE)      # error shrinks ~60x per sweep: 1e-11 after six
    assert e[-1] == pytest.approx(FCI["synth_8o8e"][_key(symm)], abs=E_TOL)
    assert min(e) > FCI["synth_8o8e"][_key(symm)] - E_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_gpu_config1_system_reaches_the_fci_energy(harness_gpu, fcidump_8o8e, symm):
    """BASELINE configs[0]'s system (8e/8o) at the full bond dimension 4^4 = 256: the GPU engine's two-site sweeps land on the
    numpy FCI energy -- an answer that no code of this repository's C++ side produced"""
    e, info = harness_gpu.ts_dmrg(fcidump_8o8e, symm, 8, 8, 16, 256, 6, GPU)
    assert e[-1] == pytest.approx(FCI["synth_8o8e"][_key(symm)], abs=E_TOL)
    assert min(e) > FCI["synth_8o8e"][_key(symm)] - E_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_gpu_full_bond_dimension_sweeps_reach_the_fci_energy_6o6e(harness_gpu, symm):
    e, info = harness_gpu.ts_dmrg("synth_6o6e.fcidump", symm, 6, 6, 8, 64, 3, GPU)
    assert e[-1] == pytest.approx(FCI["synth_6o6e"][_key(symm)], abs=E_TOL)
