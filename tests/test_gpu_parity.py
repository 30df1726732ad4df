"""Parity tests proper: the B200 path (qcm::GpuEngine -> C ABI of include/qcm_b200.h -> sm_100a kernels) against the
CPU oracle on the same seeded inputs.  Bar (BASELINE.json north_star): symmetry block structure and quantum-number
indexing bit-exact, sigma vectors and boundaries within 1e-10 relative, energies within 1e-8 Eh."""
import ctypes, json, os, tempfile
import pytest
import torch
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
TOL = 1e-10       # north_star: "sigma vectors within 1e-10 relative"
E_TOL = 1e-8      # north_star: "energies ... within 1e-8 Eh"
REF = json.load(open(os.path.join(GOLDEN, "reference_values.json")))
GPU = 1


@pytest.mark.parametrize("f,L,ne", [("synth_4o4e.fcidump", 4, 4), ("synth_6o6e.fcidump", 6, 6), ("lih_4o.fcidump", 4, 2), ("benzene_6o.fcidump", 6, 6)])
@pytest.mark.parametrize("symm", ["2u1", "su2u1", "2u1pg", "su2u1pg"])
def test_chain_parity(harness_gpu, f, L, ne, symm):
    out = harness_gpu.chain_parity(f, symm, L, ne, 20, engine=GPU)
    assert out[0] == 2 * (L + 1) and out[3] == L and out[6] == L - 1
    assert out[1] == 1 and out[4] == 1 and out[7] == 1, "block structure differs from the oracle"
    assert out[2] < TOL and out[5] < TOL and out[8] < TOL, out[:9]
    assert abs(out[10] - out[11]) < E_TOL
    assert out[9] < E_TOL            # <psi|sigma> == boundary-chain energy (test_siteproblem.cpp:38-95)


@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
def test_waves_under_a_small_workspace_budget(harness_gpu, symm):
    out = harness_gpu.chain_parity("synth_6o6e.fcidump", symm, 6, 6, 20, engine=GPU, budget=300)
    assert out[1] == 1 and out[4] == 1 and out[7] == 1
    assert out[2] < TOL and out[5] < TOL and out[8] < TOL, out[:9]


@pytest.mark.parametrize("M", [1, 3, 7])
@pytest.mark.parametrize("symm", ["2u1", "su2u1"])
def test_ragged_and_degenerate_blocks(harness_gpu, symm, M):
    out = harness_gpu.chain_parity("synth_6o6e.fcidump", symm, 6, 6, M, seed=5, engine=GPU)
    assert out[1] == 1 and out[4] == 1 and out[7] == 1
    assert out[2] < TOL and out[5] < TOL and out[8] < TOL, out[:9]


@pytest.mark.parametrize("symm", ["su2u1pg", "su2u1", "2u1pg", "2u1"])
def test_h2_energy(harness_gpu, symm):
    e, asym = harness_gpu.exact_energy("h2_2o.fcidump", symm, 2, 2, GPU)       # dmrg/tests/test1.cpp:93
    assert e == pytest.approx(REF["energies"]["h2_2o"]["value"], abs=E_TOL)


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_lih_energy(harness_gpu, symm):
    e, asym = harness_gpu.exact_energy("lih_4o.fcidump", symm, 4, 2, GPU)      # LiHFixture.h:112
    assert e == pytest.approx(REF["energies"]["lih_4o"]["value"], abs=E_TOL)
    assert asym < 1e-9


def test_h2_4o_energy(harness_gpu):
    e, _ = harness_gpu.exact_energy("h2_4o.fcidump", "su2u1pg", 4, 2, GPU)     # H2_2e4o.TI.SS.out:70
    assert e == pytest.approx(REF["energies"]["h2_4o"]["value"], abs=E_TOL)


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
@pytest.mark.parametrize("site,twosite", [(3, True), (0, True), (6, True), (4, False), (1, False)])
def test_config1_sized_site_problems(harness_gpu, fcidump_8o8e, symm, site, twosite):
    # BASELINE.json configs[0]: 8e/8o, M = 256 (sigma two-site; single-site problems also run both boundary steps)
    out = harness_gpu.synth_parity(fcidump_8o8e, symm, 8, 8, site, twosite, 256, engine=GPU)
    assert out[0] == 1 and out[1] < TOL and out[2] > 0, out[:4]
    if not twosite:
        assert out[4] == 1 and out[6] == 1 and out[5] < TOL and out[7] < TOL, out[4:8]


@pytest.mark.parametrize("symm,M", [("su2u1", 400), ("2u1", 600)])
def test_mid_sized_blocks_use_every_tile_variant(harness_gpu, symm, M):
    # 12 orbitals: sector sizes from 1 to > 128, several K-segments per output block, split-K with atomics
    from qcmaquis_b200.fcidump import make_fcidump
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "synth_12o12e.fcidump")
    make_fcidump(path, 12, 12)
    out = harness_gpu.synth_parity(path, symm, 12, 12, 5, True, M, engine=GPU)
    assert out[0] == 1 and out[1] < TOL, out[:4]
    out = harness_gpu.synth_parity(path, symm, 12, 12, 6, False, M, engine=GPU)
    assert out[0] == 1 and out[1] < TOL and out[4] == 1 and out[6] == 1 and out[5] < TOL and out[7] < TOL, out[:8]


class Driver:
    """qcmaquis_b200/lib/libqcm_host.so: the host driver bench.py uses (plan once, sigma per eigensolver iteration)."""

    def __init__(self, built, norb, nelec, symm, M, site, seed=1):
        from qcmaquis_b200.fcidump import make_fcidump
        self.cu = ctypes.CDLL(built["cuda"], mode=ctypes.RTLD_GLOBAL)
        self.cu.qcm_last_error.restype = ctypes.c_char_p
        self.host = ctypes.CDLL(built["host"])
        self.host.qcmd_create.restype = ctypes.c_void_p
        assert self.cu.qcm_init(0) == 0, self.cu.qcm_last_error()
        self.path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "s.fcidump")
        make_fcidump(self.path, norb, nelec)
        self.err = ctypes.create_string_buffer(1024)
        h = self.host.qcmd_create(self.path.encode(), symm.encode(), norb, nelec, self.err, 1024)
        assert h, self.err.value
        self.h = ctypes.c_void_p(h)
        self.info = (ctypes.c_double * 32)()
        assert self.host.qcmd_setup_site(self.h, site, 1, M, seed, 0, 0, 1, self.info, self.err, 1024) == 0, self.err.value
        self.n_psi, self.n_sigma = int(self.info[5]), int(self.info[6])

    def sigma(self, x):
        y = torch.empty(self.n_sigma, dtype=torch.float64)
        assert self.host.qcmd_sigma_host(self.h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), self.err, 1024) == 0, self.err.value
        return y

    def sigma_dev(self):
        assert self.host.qcmd_sigma_dev(self.h, 1, self.err, 1024) == 0, self.err.value
        y = torch.empty(self.n_sigma, dtype=torch.float64)
        assert self.host.qcmd_get_sigma_dev(self.h, ctypes.c_void_p(y.data_ptr())) == 0
        return y

    def psi(self):
        x = torch.empty(self.n_psi, dtype=torch.float64)
        self.host.qcmd_get_psi(self.h, ctypes.c_void_p(x.data_ptr()))
        return x


@pytest.fixture(scope="module")
def driver_cfg2(built):
    return Driver(built, 26, 10, "su2u1", 1000, 12)      # BASELINE.json configs[1]


def test_full_size_linearity(driver_cfg2):
    # size-independent property at BASELINE's full size: sigma is linear in psi
    d = driver_cfg2
    g = torch.Generator().manual_seed(7)
    x = torch.randn(d.n_psi, dtype=torch.float64, generator=g); y = torch.randn(d.n_psi, dtype=torch.float64, generator=g)
    hx, hy, hz = d.sigma(x), d.sigma(y), d.sigma(0.75 * x - 1.5 * y)
    ref = 0.75 * hx - 1.5 * hy
    assert float((hz - ref).norm() / ref.norm()) < 1e-12
    assert float(d.sigma(torch.zeros(d.n_psi, dtype=torch.float64)).abs().max()) == 0.0


def test_full_size_host_and_device_paths_agree(driver_cfg2):
    d = driver_cfg2
    a, b = d.sigma(d.psi()), d.sigma_dev()
    assert float((a - b).norm() / a.norm()) < 1e-13


def test_full_size_against_the_oracle(built, driver_cfg2):
    d = driver_cfg2
    olib = ctypes.CDLL(built["oracle"])
    olib.orc_create.restype = ctypes.c_void_p
    err = ctypes.create_string_buffer(1024)
    oh = ctypes.c_void_p(olib.orc_create(d.path.encode(), b"su2u1", 26, 10, err, 1024))
    assert oh.value, err.value
    pe, sec, ov, se = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    assert olib.orc_setup_site(oh, 12, 1, 1000, 1, ctypes.byref(pe), err, 1024) == 0, err.value
    assert olib.orc_sigma(oh, 1, ctypes.byref(sec), ctypes.byref(ov), ctypes.byref(se), err, 1024) == 0, err.value
    assert int(se.value) == d.n_sigma, "sigma block structure differs from the oracle"
    ref = torch.empty(d.n_sigma, dtype=torch.float64)
    olib.orc_get_sigma(oh, ctypes.c_void_p(ref.data_ptr()))
    got = d.sigma(d.psi())
    assert float((got - ref).norm() / ref.norm()) < TOL
    olib.orc_destroy(oh)


def test_vector_algebra(built):
    # solver-side BLAS-1 on device arrays (mpstensor.hpp:346-395,458-522)
    cu = ctypes.CDLL(built["cuda"])
    cu.qcm_last_error.restype = ctypes.c_char_p
    assert cu.qcm_init(0) == 0, cu.qcm_last_error()
    n = 1_000_003
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, dtype=torch.float64, generator=g); y = torch.randn(n, dtype=torch.float64, generator=g)
    ax, ay = ctypes.c_void_p(), ctypes.c_void_p()
    assert cu.qcm_array_alloc(ctypes.c_int64(n), ctypes.byref(ax)) == 0 and cu.qcm_array_alloc(ctypes.c_int64(n), ctypes.byref(ay)) == 0
    cu.qcm_array_upload(ax, ctypes.c_int64(0), ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(n))
    cu.qcm_array_upload(ay, ctypes.c_int64(0), ctypes.c_void_p(y.data_ptr()), ctypes.c_int64(n))
    r = ctypes.c_double()
    assert cu.qcm_vec_dot(ax, ay, ctypes.c_int64(n), ctypes.byref(r)) == 0
    assert r.value == pytest.approx(float(x @ y), rel=1e-12)
    assert cu.qcm_vec_axpy(ctypes.c_double(-0.5), ax, ay, ctypes.c_int64(n)) == 0
    assert cu.qcm_vec_scal(ctypes.c_double(3.0), ay, ctypes.c_int64(n)) == 0
    out = torch.empty(n, dtype=torch.float64)
    cu.qcm_array_download(ay, ctypes.c_int64(0), ctypes.c_void_p(out.data_ptr()), ctypes.c_int64(n))
    assert torch.equal(out, 3.0 * torch.addcmul(y, torch.full_like(x, -0.5), x)) or float((out - 3.0 * (y - 0.5 * x)).abs().max()) < 1e-14
    # range errors are reported, not ignored
    assert cu.qcm_vec_dot(ax, ay, ctypes.c_int64(n + 1), ctypes.byref(r)) != 0
    cu.qcm_array_free(ax); cu.qcm_array_free(ay)


@pytest.mark.parametrize("symm,norb,nelec,M", [("su2u1", 8, 8, 64), ("2u1", 8, 8, 48), ("su2u1", 12, 12, 100)])
def test_synthetic_start_sweep_matches_the_oracle(built, symm, norb, nelec, M):
    """The sweep leg of bench.py (qcmd_ts_sweeps_synth: two-site sweep from a random MPS on the synthetic sector lists, stale
    boundaries dropped, device-resident Jacobi-Davidson) against the same driver on the CPU oracle: every micro-iteration
    energy within 1e-8 Eh."""
    from qcmaquis_b200.fcidump import make_fcidump
    cu = ctypes.CDLL(built["cuda"], mode=ctypes.RTLD_GLOBAL)
    cu.qcm_last_error.restype = ctypes.c_char_p
    assert cu.qcm_init(0) == 0, cu.qcm_last_error()
    host = ctypes.CDLL(built["host"]); host.qcmd_create.restype = ctypes.c_void_p
    olib = ctypes.CDLL(built["oracle"]); olib.orc_create.restype = ctypes.c_void_p
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "s.fcidump")
    make_fcidump(path, norb, nelec)
    err = ctypes.create_string_buffer(1024)
    h = ctypes.c_void_p(host.qcmd_create(path.encode(), symm.encode(), norb, nelec, err, 1024)); assert h.value, err.value
    oh = ctypes.c_void_p(olib.orc_create(path.encode(), symm.encode(), norb, nelec, err, 1024)); assert oh.value, err.value
    eg = (ctypes.c_double * 4096)(); ng = ctypes.c_int(); ig = (ctypes.c_double * 32)()
    eo = (ctypes.c_double * 4096)(); no = ctypes.c_int(); io = (ctypes.c_double * 32)()
    host.qcmd_ts_sweeps_synth.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    assert host.qcmd_ts_sweeps_synth(h, M, 2, 3, 0, 0, 1, 0, 0.0, eg, 4096, ctypes.byref(ng), ig, err, 1024) == 0, err.value
    assert olib.orc_ts_sweeps_synth(oh, M, 2, 3, 0, eo, 4096, ctypes.byref(no), io, err, 1024) == 0, err.value
    assert ng.value == no.value == 2 * (2 * norb - 2)
    assert max(abs(eg[i] - eo[i]) for i in range(ng.value)) < 1e-8
    assert int(ig[0]) == int(io[0]), "different numbers of sigma evaluations"
    assert ig[16] > 0 and ig[17] > 0      # the engine booked its FLOPs
    host.qcmd_destroy(h); olib.orc_destroy(oh)


def test_sharded_plan_needs_a_matching_communicator(built):
    """A plan built for rank r of N produces a partial result: executing it without an N-rank communicator must fail loudly
    (the sharding lives in the plan and is checked against qcm_comm_init's state)."""
    from qcmaquis_b200.fcidump import make_fcidump
    cu = ctypes.CDLL(built["cuda"], mode=ctypes.RTLD_GLOBAL)
    cu.qcm_last_error.restype = ctypes.c_char_p
    assert cu.qcm_init(0) == 0, cu.qcm_last_error()
    host = ctypes.CDLL(built["host"]); host.qcmd_create.restype = ctypes.c_void_p
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "s.fcidump")
    make_fcidump(path, 8, 8)
    err = ctypes.create_string_buffer(1024)
    h = ctypes.c_void_p(host.qcmd_create(path.encode(), b"su2u1", 8, 8, err, 1024)); assert h.value, err.value
    info = (ctypes.c_double * 32)()
    assert host.qcmd_setup_site(h, 3, 1, 64, 1, 0, 1, 2, info, err, 1024) == 0, err.value      # rank 1 of 2, no communicator
    assert host.qcmd_sigma_dev(h, 1, err, 1024) != 0
    assert b"communicator" in err.value
    host.qcmd_destroy(h)


def test_spill_tier_keeps_the_sweep_unchanged(built, monkeypatch):
    """Boundaries the sweep leaves behind are evicted to pinned host memory and prefetched before they are needed again
    (qcm_array_evict / qcm_array_prefetch, the storage::disk protocol of utils/storage.h:113-185 around every site): same
    energies as with everything resident, and the tier is really used."""
    from qcmaquis_b200.fcidump import make_fcidump
    cu = ctypes.CDLL(built["cuda"], mode=ctypes.RTLD_GLOBAL)
    cu.qcm_last_error.restype = ctypes.c_char_p
    assert cu.qcm_init(0) == 0, cu.qcm_last_error()
    host = ctypes.CDLL(built["host"]); host.qcmd_create.restype = ctypes.c_void_p
    host.qcmd_ts_sweeps_synth.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "s.fcidump")
    make_fcidump(path, 10, 10)
    err = ctypes.create_string_buffer(1024)
    runs = []
    for spill in (False, True):
        if spill:
            monkeypatch.setenv("QCM_SPILL", "1")
        else:
            monkeypatch.delenv("QCM_SPILL", raising=False)
        h = ctypes.c_void_p(host.qcmd_create(path.encode(), b"su2u1", 10, 10, err, 1024)); assert h.value, err.value
        e = (ctypes.c_double * 4096)(); n = ctypes.c_int(); info = (ctypes.c_double * 32)()
        assert host.qcmd_ts_sweeps_synth(h, 120, 2, 3, 0, 0, 1, 0, 0.0, e, 4096, ctypes.byref(n), info, err, 1024) == 0, err.value
        runs.append(([e[i] for i in range(n.value)], info[21], info[22]))
        host.qcmd_destroy(h)
    (e0, ev0, pf0), (e1, ev1, pf1) = runs
    assert len(e0) == len(e1) == 2 * 18
    assert max(abs(a - b) for a, b in zip(e0, e1)) < 1e-10
    assert ev0 == 0 and ev1 > 10 and pf1 > 10, (ev0, ev1, pf1)

    # the C entry points themselves: evict -> not usable -> prefetch -> same contents
    n_el = 1 << 20
    x = torch.randn(n_el, dtype=torch.float64)
    a = ctypes.c_void_p(); assert cu.qcm_array_alloc(ctypes.c_int64(n_el), ctypes.byref(a)) == 0
    assert cu.qcm_array_upload(a, ctypes.c_int64(0), ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(n_el)) == 0
    pin = ctypes.c_void_p(); assert cu.qcm_pinned_alloc(ctypes.c_int64(n_el), ctypes.byref(pin)) == 0
    assert cu.qcm_array_evict(a, pin) == 0
    r = ctypes.c_double()
    assert cu.qcm_vec_dot(a, a, ctypes.c_int64(n_el), ctypes.byref(r)) != 0 and b"evicted" in cu.qcm_last_error()
    assert cu.qcm_array_prefetch(a, pin) == 0
    assert cu.qcm_vec_dot(a, a, ctypes.c_int64(n_el), ctypes.byref(r)) == 0
    assert r.value == pytest.approx(float(x @ x), rel=1e-12)
    cu.qcm_array_free(a); cu.qcm_pinned_free(pin)


@pytest.mark.parametrize("symm,M", [("su2u1", 120), ("2u1", 60)])
@pytest.mark.parametrize("slices", [2, 3])
def test_time_sliced_shards_match_the_oracle(harness_gpu, monkeypatch, symm, M, slices):
    """qcm_site_hamil2_sliced: the site problem planned as `slices` shards (the plans that many ranks would run) and executed one
    after another on ONE device -- the single-GPU mode for problems whose resident step-1 products exceed the device (cfg4).
    Same sigma as the oracle, structure included."""
    from qcmaquis_b200.fcidump import make_fcidump
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "synth_12o12e.fcidump")
    make_fcidump(path, 12, 12)
    monkeypatch.setenv("QCM_SLICES", str(slices))
    out = harness_gpu.synth_parity(path.encode(), symm, 12, 12, 5, True, M, engine=GPU)
    assert out[0] == 1 and out[1] < TOL, out[:4]


def test_store_mode_output_with_a_long_k_list(harness_gpu):
    """Task-array level of the C ABI: a step-1 (store-mode) output whose K-segment list exceeds 96 K-chunks must run in one work
    item (an SU2 site whose leading sector passes ~1.5k rows), an accumulating output of the same length is split and combined
    with FP64 atomics -- both against host dgemm."""
    out = (ctypes.c_double * 2)(); err = ctypes.create_string_buffer(1024)
    assert harness_gpu.lib.qcmt_long_k_plan(out, err, 1024) == 0, err.value.decode()
    assert out[0] < 1e-12 and out[1] > 1.0, list(out)
