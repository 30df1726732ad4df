"""Pins the CPU oracle (oracle/oracle_engine.hpp) and the host data model it runs on against the reference's own
known answers (tests/golden/reference_values.json, produced by tests/golden/make_goldens.py from the reference
tree): Wigner 9j table, MPO bond dimensions / Hermitian pairs / term counts printed by shipped example runs,
ground-state energies of the reference's end-to-end tests, and the sigma / boundary energy identities its
hot-path tests check (test_siteproblem.cpp:38-95, BoundaryPropagatorElectronic.cpp:39-66)."""
import json, os
import pytest
from conftest import GOLDEN

REF = json.load(open(os.path.join(GOLDEN, "reference_values.json")))
ORACLE, INTERP = 2, 0


def test_wigner_9j_table(harness_cpu):
    # dmrg/tests/test_wigner.cpp:21-46, BOOST_CHECK_CLOSE 1e-6 percent
    for args, want in REF["wigner_9j"]["cases"]:
        got = harness_cpu.lib.qcmt_wigner9j(*args)
        assert got == pytest.approx(want, rel=1e-8, abs=1e-9), (args, got, want)


def test_wigner_6j_known_values(harness_cpu):
    # {1/2 1/2 1; 1/2 1/2 0} = 1/2,  {1 1 1; 1 1 1} = 1/6,  {1/2 1/2 0; 1/2 1/2 0} = -1/2 (arguments 2j)
    w = harness_cpu.lib.qcmt_wigner6j
    assert w(1, 1, 2, 1, 1, 0) == pytest.approx(0.5, abs=1e-14)
    assert w(2, 2, 2, 2, 2, 2) == pytest.approx(1.0 / 6.0, abs=1e-14)
    assert w(1, 1, 0, 1, 1, 0) == pytest.approx(-0.5, abs=1e-14)


@pytest.mark.parametrize("key,f,L,ne", [("h2_4o/2u1pg", "h2_4o.fcidump", 4, 2), ("h2_4o/su2u1pg", "h2_4o.fcidump", 4, 2),
                                        ("benzene_6o/su2u1pg", "benzene_6o.fcidump", 6, 6)])
def test_mpo_bond_dimensions_and_hermitian_pairs(harness_cpu, key, f, L, ne):
    symm = key.split("/")[1]
    dims, pairs, nterms, _ = harness_cpu.mpo_dims(f, symm, L, ne)
    assert dims == REF["mpo"][key]["dims"]          # bit-exact bond indexing: "MPO Bond p: dim/pairs"
    assert pairs == REF["mpo"][key]["pairs"]
    assert nterms == REF["mpo"][key]["terms"]       # "The hamiltonian will contain N terms"


def test_dense_integral_mpo_law(harness_cpu):
    # B_SU2(l) = 4 l^2 + 2 l + 8 r + 2 for dense integral tables (r >= 2), 10 at r = 1, 1 at the end
    dims, _, _, _ = harness_cpu.mpo_dims("synth_6o6e.fcidump", "su2u1", 6, 6)
    want = [4 * l * l + 2 * l + 8 * (6 - l) + 2 for l in range(1, 5)] + [10, 1]
    assert dims == want


@pytest.mark.parametrize("symm", ["su2u1pg", "su2u1", "2u1pg", "2u1"])
@pytest.mark.parametrize("engine", [ORACLE, INTERP])
def test_h2_energy(harness_cpu, symm, engine):
    # dmrg/tests/test1.cpp:93 -- same value for all four symmetry groups
    e, asym = harness_cpu.exact_energy("h2_2o.fcidump", symm, 2, 2, engine)
    assert e == pytest.approx(REF["energies"]["h2_2o"]["value"], abs=1e-10)
    assert asym < 1e-12


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
@pytest.mark.parametrize("engine", [ORACLE, INTERP])
def test_lih_energy(harness_cpu, symm, engine):
    # dmrg/tests/Fixtures/LiHFixture.h:112 (DMRG at m=100 on 4 orbitals is exact); two-site centre on complete bases
    e, asym = harness_cpu.exact_energy("lih_4o.fcidump", symm, 4, 2, engine)
    assert e == pytest.approx(REF["energies"]["lih_4o"]["value"], abs=1e-8)
    assert asym < 1e-10


@pytest.mark.parametrize("symm", ["su2u1pg", "2u1pg"])
def test_h2_4o_energy(harness_cpu, symm):
    # examples/iTD-DMRG/H2_2e4o.TI.SS.out:70 -- converged energy of the shipped example run
    e, _ = harness_cpu.exact_energy("h2_4o.fcidump", symm, 4, 2, ORACLE)
    assert e == pytest.approx(REF["energies"]["h2_4o"]["value"], abs=1e-8)


@pytest.mark.parametrize("f,L,ne", [("lih_4o.fcidump", 4, 2), ("benzene_6o.fcidump", 6, 6)])
@pytest.mark.parametrize("symm", ["su2u1pg", "2u1pg"])
def test_sigma_and_boundary_energy_identity(harness_cpu, f, L, ne, symm):
    # test_siteproblem.cpp:38-95: <psi|site_hamil2(psi)> equals the expectation value from the boundary chain on
    # every site; BoundaryPropagatorElectronic.cpp:39-66: the last left boundary has one 1x1 block whose trace is it.
    # engine = oracle is checked through the interpreter call too (out[9] uses the engine's own boundaries)
    out = harness_cpu.chain_parity(f, symm, L, ne, 16, seed=3, engine=INTERP)
    assert out[1] == 1 and out[4] == 1 and out[7] == 1
    assert abs(out[10] - out[11]) < 1e-12 * max(1.0, abs(out[11]))
    assert out[9] < 1e-11 * max(1.0, abs(out[11]))
    if symm.startswith("su2"):
        assert out[13] < 1e-12     # the oracle's lbtm and rbtm SU2 variants agree (non-abelian/site_hamil.hpp:35-38)
