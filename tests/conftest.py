"""Shared fixtures.  CPU tests (`-m "not gpu"`) use the oracle, the host logic and the plan interpreter; GPU tests
(`-m gpu`) go through the C ABI of include/qcm_b200.h and compare with the oracle on the same seeded inputs."""
import ctypes, os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); runs through the C ABI library")


def golden(name):
    if isinstance(name, bytes):
        return name
    return os.path.join(GOLDEN, name).encode()


class Harness:
    """ctypes view of tests/harness/libqcm_harness*.so (oracle + engine under test)."""

    def __init__(self, lib):
        self.lib = lib
        lib.qcmt_wigner9j.restype = ctypes.c_double
        lib.qcmt_wigner6j.restype = ctypes.c_double

    def chain_parity(self, f, symm, L, ne, M, seed=42, engine=0, world=1, budget=1 << 28):
        out = (ctypes.c_double * 16)(); err = ctypes.create_string_buffer(1024)
        rc = self.lib.qcmt_chain_parity(golden(f), symm.encode(), L, ne, M, seed, engine, world, ctypes.c_longlong(budget), out, 16, err, 1024)
        assert rc == 0, err.value.decode()
        return list(out)

    def synth_parity(self, f, symm, L, ne, site, twosite, M, seed=1, engine=0, world=1, budget=1 << 28):
        out = (ctypes.c_double * 16)(); err = ctypes.create_string_buffer(1024)
        rc = self.lib.qcmt_synth_parity(golden(f), symm.encode(), L, ne, site, int(twosite), M, seed, engine, world, ctypes.c_longlong(budget), out, 16, err, 1024)
        assert rc == 0, err.value.decode()
        return list(out)

    def exact_energy(self, f, symm, L, ne, engine):
        e, a = ctypes.c_double(), ctypes.c_double(); err = ctypes.create_string_buffer(1024)
        rc = self.lib.qcmt_exact_energy(golden(f), symm.encode(), L, ne, engine, ctypes.byref(e), ctypes.byref(a), err, 1024)
        assert rc == 0, err.value.decode()
        return e.value, a.value

    def ss_dmrg(self, f, symm, L, ne, M, nsweeps, engine, seed=42):
        """single-site DMRG sweeps (qcm/sweep.hpp) -> (energies per micro-iteration, info)"""
        e = (ctypes.c_double * 8192)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
        rc = self.lib.qcmt_ss_dmrg(golden(f), symm.encode(), L, ne, M, nsweeps, seed, engine, e, 8192, ctypes.byref(n), info, err, 1024)
        assert rc == 0, err.value.decode()
        return list(e[:n.value]), list(info)

    def ts_dmrg(self, f, symm, L, ne, M0, M, nsweeps, engine, seed=42):
        """two-site DMRG sweeps (qcm/twosite.hpp) from a random MPS of bond dimension M0, truncation to M"""
        e = (ctypes.c_double * 8192)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
        rc = self.lib.qcmt_ts_dmrg(golden(f), symm.encode(), L, ne, M0, M, nsweeps, seed, engine, e, 8192, ctypes.byref(n), info, err, 1024)
        assert rc == 0, err.value.decode()
        return list(e[:n.value]), list(info)

    def hdiag_parity(self, f, symm, L, ne, M, engine, seed=42):
        out = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
        rc = self.lib.qcmt_hdiag_parity(golden(f), symm.encode(), L, ne, M, seed, engine, out, err, 1024)
        assert rc == 0, err.value.decode()
        return list(out)

    def mpo_dims(self, f, symm, L, ne):
        dims, pairs = (ctypes.c_int * L)(), (ctypes.c_int * L)(); core = ctypes.c_double(); n = ctypes.c_int(); err = ctypes.create_string_buffer(1024)
        assert self.lib.qcmt_mpo_dims(golden(f), symm.encode(), L, ne, dims, pairs, ctypes.byref(core), err, 1024) == 0, err.value.decode()
        assert self.lib.qcmt_num_terms(golden(f), symm.encode(), L, ne, ctypes.byref(n), err, 1024) == 0, err.value.decode()
        return list(dims), list(pairs), n.value, core.value


@pytest.fixture(scope="session")
def built():
    from qcmaquis_b200 import build
    return build.build_all()


@pytest.fixture(scope="session")
def harness_cpu_path():
    """the CPU harness (oracle + plan interpreter) needs g++ and OpenBLAS only -- no CUDA toolkit"""
    from qcmaquis_b200 import build
    return build.build_harness(gpu=False)[0]


@pytest.fixture(scope="session")
def harness_cpu(harness_cpu_path):
    return Harness(ctypes.CDLL(harness_cpu_path))


@pytest.fixture(scope="session")
def harness_gpu(built):
    h = Harness(ctypes.CDLL(built["harness"][1]))
    if h.lib.qcmt_gpu_available() < 1:
        pytest.fail("a test marked gpu ran without a CUDA device: the hot path has no CPU fallback")
    return h


@pytest.fixture(scope="session")
def fcidump_8o8e():
    """BASELINE config 1 system: 8 electrons in 8 orbitals, the deterministic synthetic integrals of SURVEY 8(d)"""
    import tempfile
    from qcmaquis_b200.fcidump import make_fcidump
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "synth_8o8e.fcidump")
    make_fcidump(path, 8, 8)
    return path.encode()
