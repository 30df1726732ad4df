"""Host logic: the flattened schedule (qcmaquis_b200/csrc/qcm/plan.hpp) executed by a plain-loop interpreter must
reproduce the oracle bit-for-bit in block structure and to rounding in values -- boundaries along whole chains,
single-site and two-site sigma -- for all four symmetry groups, with the workspace budget forcing several waves,
and with the MPO bond index sharded over 2 and 3 ranks (partial results summed, as the allreduce does)."""
import pytest

TOL = 1e-12   # same arithmetic, different summation order


@pytest.mark.parametrize("f,L,ne", [("synth_4o4e.fcidump", 4, 4), ("synth_6o6e.fcidump", 6, 6)])
@pytest.mark.parametrize("symm", ["2u1", "su2u1", "2u1pg", "su2u1pg"])
def test_chain_parity(harness_cpu, f, L, ne, symm):
    out = harness_cpu.chain_parity(f, symm, L, ne, 20)
    assert out[0] == 2 * (L + 1) and out[3] == L and out[6] == L - 1
    assert out[1] == 1 and out[4] == 1 and out[7] == 1, "block structure differs from the oracle"
    assert out[2] < TOL and out[5] < TOL and out[8] < TOL, out[:9]
    assert out[9] < 1e-10 * max(1.0, abs(out[11]))


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("budget", [300, 5000])
def test_small_workspace_budget_forces_waves(harness_cpu, symm, budget):
    out = harness_cpu.chain_parity("synth_6o6e.fcidump", symm, 6, 6, 20, budget=budget)
    assert out[1] == 1 and out[4] == 1 and out[7] == 1
    assert out[2] < TOL and out[5] < TOL and out[8] < TOL, out[:9]


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_plans_sum_to_the_full_result(harness_cpu, symm, world):
    out = harness_cpu.chain_parity("synth_6o6e.fcidump", symm, 6, 6, 20, world=world)
    assert out[1] == 1 and out[4] == 1 and out[7] == 1
    assert out[2] < TOL and out[5] < TOL and out[8] < TOL, out[:9]


def test_real_integrals_with_missing_terms(harness_cpu):
    # the LiH table has symmetry-zero integrals -> fewer bond labels than the dense law; benzene is dense
    for f, L, ne in [("lih_4o.fcidump", 4, 2), ("benzene_6o.fcidump", 6, 6)]:
        for symm in ["su2u1", "2u1"]:
            out = harness_cpu.chain_parity(f, symm, L, ne, 12, seed=11)
            assert out[1] == 1 and out[4] == 1 and out[7] == 1
            assert out[2] < TOL and out[5] < TOL and out[8] < TOL, (f, symm, out[:9])


@pytest.mark.parametrize("symm,twosite", [("su2u1", True), ("2u1", True), ("su2u1", False), ("2u1", False)])
def test_synthetic_site_generator(harness_cpu, symm, twosite):
    # the generator behind bench.py (fabricated sectors and boundaries around the true MPO), small M
    out = harness_cpu.synth_parity("synth_6o6e.fcidump", symm, 6, 6, 2, twosite, 24)
    assert out[0] == 1 and out[1] < TOL and out[2] > 0
    if not twosite:
        assert out[4] == 1 and out[6] == 1 and out[5] < TOL and out[7] < TOL


def test_ragged_and_degenerate_inputs(harness_cpu):
    # M = 1: every sector has size 1 (all GEMMs degenerate to scalars); M = 3 with 6 orbitals: ragged 1..2 blocks
    for M in (1, 3):
        for symm in ("su2u1", "2u1"):
            out = harness_cpu.chain_parity("synth_6o6e.fcidump", symm, 6, 6, M, seed=5)
            assert out[1] == 1 and out[4] == 1 and out[7] == 1
            assert out[2] < TOL and out[5] < TOL and out[8] < TOL


@pytest.mark.parametrize("world", [2, 4, 8])
def test_edge_sharding_computes_no_step1_product_twice(harness_cpu, world):
    """the edges of the MPO bond graph are sharded by their step-1 index -- every T[b] is computed by exactly one
    rank, the algorithmic FLOP count is booked exactly once, and the heaviest rank stays close to its fair share."""
    import ctypes, os, tempfile
    from qcmaquis_b200.fcidump import make_fcidump
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "synth_12o12e.fcidump")
    make_fcidump(path, 12, 12)
    out = (ctypes.c_double * 4)(); err = ctypes.create_string_buffer(1024)
    rc = harness_cpu.lib.qcmt_shard_stats(path.encode(), b"su2u1", 12, 12, 5, 300, 1, world, out, err, 1024)
    assert rc == 0, err.value.decode()
    assert out[0] == pytest.approx(1.0, abs=1e-12)       # step 1 is partitioned, not replicated
    assert out[2] == pytest.approx(1.0, abs=1e-12)       # FLOPs booked once across ranks
    assert 1.0 - 1e-12 <= out[3] < 1.03                   # no closing product is repeated: partial W sums are exchanged (reduce-scatter); the
                                                         # excess is row units of exchanged blocks that no bond feeds (closed as zeros)
    assert out[1] < 1.25                                 # the heaviest rank stays close to its fair share


@pytest.mark.parametrize("symm,M", [("su2u1", 100), ("2u1", 60)])
@pytest.mark.parametrize("world", [1, 3])
def test_twelve_orbital_two_site_plan(harness_cpu, symm, M, world):
    """A problem large enough for the W pass to form its 32/64-destination classes: exercises the piggy-back pass of the
    grouping (small high fan-in groups moved into idle columns of large ones) and, for world > 1, an exchange region of
    several hundred thousand elements."""
    import os, tempfile
    from qcmaquis_b200.fcidump import make_fcidump
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "synth_12o12e.fcidump")
    make_fcidump(path, 12, 12)
    out = harness_cpu.synth_parity(path.encode(), symm, 12, 12, 5, True, M, engine=0, world=world)
    assert out[0] == 1 and out[1] < TOL, out[:4]
    if world > 1:
        assert out[10] > 1e5
