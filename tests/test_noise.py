"""Noise-perturbed truncation (SURVEY 8(a9) remainder): the contractions behind predict_new_state_l2r/r2l_sweep
(contractions/common/prediction.hpp:19-57, 84-130 -> move_boundary.hpp:68-126 left/right_boundary_tensor_mpo) and
TwoSiteTensor::predict_split_l2r / r2l (mp_tensors/twositetensor.hpp:184-300) -- the reference's DEFAULT split path
(alpha_initial = 1e-2, utils/DmrgParameters.h:47-49; ts_optimize.hpp:198-215).  Engine under test (plan interpreter on CPU,
qcm::GpuEngine through the C ABI on the GPU) against the oracle's restatement; the reference's own printed energies and
Jacobi-Davidson iteration counts of examples/iTD-DMRG/H2_2e4o.TI.SS.out pin the oracle."""
import ctypes, json, os
import pytest
from conftest import golden, GOLDEN


def _noise_parity(h, f, symm, L, ne, M, engine, budget=1 << 40, seed=3):
    out = (ctypes.c_double * 4)(); err = ctypes.create_string_buffer(1024)
    rc = h.lib.qcmt_noise_parity(golden(f), symm.encode(), L, ne, M, seed, engine, ctypes.c_longlong(budget), out, err, 1024)
    assert rc == 0, err.value.decode()
    return list(out)


def _check_noise_parity(h, symm, engine, budget=1 << 40):
    n, st, diff, norm = _noise_parity(h, "synth_6o6e.fcidump", symm, 6, 6, 12, engine, budget)
    assert n == 22            # 6 sites x (left, right) + 5 bonds x (fat right index, fat left index)
    assert st == 1, "block structure of the kept noise blocks differs from the oracle's"
    assert diff < 1e-12, diff
    assert norm > 1e-3        # the comparison is not between zeros


@pytest.mark.parametrize("symm", ["2u1", "su2u1", "2u1pg", "su2u1pg"])
def test_noise_term_plan_interpreter_vs_oracle(harness_cpu, symm):
    _check_noise_parity(harness_cpu, symm, 0)


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
def test_noise_term_multi_wave_plan_vs_oracle(harness_cpu, symm):
    """tiny workspace budget: the noise plan is cut into several waves"""
    _check_noise_parity(harness_cpu, symm, 0, budget=2000)


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["2u1", "su2u1", "2u1pg", "su2u1pg"])
def test_noise_term_gpu_vs_oracle(harness_gpu, symm):
    _check_noise_parity(harness_gpu, symm, 1)


@pytest.mark.gpu
def test_noise_term_gpu_multi_wave(harness_gpu):
    _check_noise_parity(harness_gpu, "su2u1", 1, budget=2000)


# ---- the reference's own run: single-site, init_type = const, alpha_initial = 1e-10, truncation 1e-50, M = 100 ----------------
def _const_noise_run(h, engine):
    e = (ctypes.c_double * 64)(); ns = (ctypes.c_int * 64)(); ba = (ctypes.c_int * 64)(); n = ctypes.c_int(); err = ctypes.create_string_buffer(1024)
    rc = h.lib.qcmt_ss_dmrg_const_noise(golden("h2_4o.fcidump"), b"2u1pg", 4, 2, 5, 1, engine, ctypes.c_double(1e-10), ctypes.c_double(1e-50), 100,
                                        e, ns, ba, 64, ctypes.byref(n), err, 1024)
    assert rc == 0, err.value.decode()
    return list(e[:n.value]), list(ns[:n.value]), list(ba[:n.value])


def _check_const_noise(energies, n_sigma, bond_after):
    """examples/iTD-DMRG/H2_2e4o.TI.SS.out:42-88 (input H2_2e4o.TI.SS.inp: single-site, init_type = const, alpha_initial = 1e-10,
    truncation 1e-50, max_bond_dimension = 100, ietl_jcd_maxiter = 10, ietl_jcd_tol = 1e-8): energies -1.129279858917138,
    -1.138235383455172, -1.15168273493923, -1.151682732118105 after 4, 10, 8, 1 Jacobi-Davidson iterations, all later
    micro-iterations converge in 1 iteration at -1.151682732118105; bond dimensions after truncation 4, 9, 14 (of 16).
    The noise term grows the bond between sites 1 and 2 from 1 back to 9 states ("Bond dimension before truncation: 9");
    without it (test_sweeps.py) the sweep stalls at a higher energy.  Energies are solver-converged to ietl_jcd_tol = 1e-8,
    which is the tolerance the north star asks for."""
    ref = json.load(open(os.path.join(GOLDEN, "reference_values.json")))["h2_4o_microiteration_energies_2u1pg_singlesite_const_init"]
    for i in range(4):
        assert abs(energies[i] - ref[i]) < 1e-8, (i, energies[i], ref[i])
    for e in energies[4:]:
        assert abs(e - ref[3]) < 1e-8
    assert n_sigma[:5] == [4, 10, 8, 1, 1], n_sigma
    assert all(k == 1 for k in n_sigma[3:]), n_sigma
    # 16 -> 14 in the reference's run: the states dropped there are eigenvalues of roundoff size (cutoff 1e-50), so only the
    # first two counts are reproducible to the digit
    assert bond_after[:2] == [4, 9], bond_after
    assert 9 <= bond_after[2] <= 16, bond_after


def test_reference_run_with_noise_oracle(harness_cpu):
    _check_const_noise(*_const_noise_run(harness_cpu, -1))


def test_reference_run_with_noise_plan_interpreter(harness_cpu):
    _check_const_noise(*_const_noise_run(harness_cpu, 0))


@pytest.mark.gpu
def test_reference_run_with_noise_gpu(harness_gpu):
    _check_const_noise(*_const_noise_run(harness_gpu, 1))


# ---- two-site sweeps with the reference's default noise schedule start (alpha = 1e-2) ---------------------------------------------
def _ts_noise(h, symm, engine, alpha, nsweeps=2):
    e = (ctypes.c_double * 512)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
    rc = h.lib.qcmt_ts_dmrg_noise(golden("synth_6o6e.fcidump"), symm.encode(), 6, 6, 4, 20, nsweeps, 42, engine, ctypes.c_double(alpha),
                                  ctypes.c_double(1e-12), e, 512, ctypes.byref(n), info, err, 1024)
    assert rc == 0, err.value.decode()
    return list(e[:n.value]), list(info)


def _check_ts_noise(h, symm, engine):
    ref, iref = _ts_noise(h, symm, -1, 1e-2)
    got, igot = _ts_noise(h, symm, engine, 1e-2)
    assert len(ref) == len(got) == 2 * (2 * 6 - 2)
    assert max(abs(a - b) for a, b in zip(ref, got)) < 1e-8
    assert iref[0] == igot[0]                  # same number of sigma evaluations
    assert igot[3] <= 21                       # truncation to M = 20 (ties at the cut may keep one more)
    plain, _ = _ts_noise(h, symm, -1, 0.)
    assert max(abs(a - b) for a, b in zip(ref[3:], plain[3:])) > 1e-6, "alpha = 1e-2 must change the sweep"


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
def test_two_site_sweeps_with_noise_plan_interpreter(harness_cpu, symm):
    _check_ts_noise(harness_cpu, symm, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
def test_two_site_sweeps_with_noise_gpu(harness_gpu, symm):
    _check_ts_noise(harness_gpu, symm, 1)
