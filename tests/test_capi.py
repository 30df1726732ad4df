"""The drop-in boundary: libqcm_b200.so loads, exports every symbol include/qcm_b200.h declares, and fails loudly
(status + message, no CPU fallback) when there is no usable device.  No compute calls here."""
import ctypes, os, re
import pytest
from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "qcm_b200.h")).read()
    return sorted(set(re.findall(r"\b(qcm_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_engine_calls():
    names = declared_symbols()
    for n in ["qcm_site_hamil2", "qcm_site_hamil2_dev", "qcm_boundary_step", "qcm_plan_create", "qcm_array_alloc", "qcm_comm_init", "qcm_last_error"]:
        assert n in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built["cuda"])
    for n in declared_symbols():
        assert hasattr(lib, n), n


def test_host_library_loads(built):
    ctypes.CDLL(built["cuda"], mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(built["host"])
    for n in ["qcmd_create", "qcmd_setup_site", "qcmd_sigma_host", "qcmd_sigma_dev", "qcmd_plan_flops"]:
        assert hasattr(lib, n), n


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = ctypes.CDLL(built["cuda"])
    lib.qcm_last_error.restype = ctypes.c_char_p
    assert lib.qcm_init(0) != 0
    assert b"CUDA" in lib.qcm_last_error() or b"device" in lib.qcm_last_error()
    arr = ctypes.c_void_p()
    assert lib.qcm_array_alloc(ctypes.c_int64(16), ctypes.byref(arr)) != 0     # every entry point refuses to run
    assert b"qcm_init" in lib.qcm_last_error()


def test_product_libraries_do_not_link_the_oracle(built):
    import subprocess
    for key in ("cuda", "host"):
        out = subprocess.run(["nm", "-D", "--defined-only", built[key]], capture_output=True, text=True).stdout
        assert "oracle" not in out and "orc_" not in out


@pytest.mark.gpu
def test_error_conventions_on_the_device(built):
    """Every call returns a status; misuse is reported through qcm_last_error(), never silently computed around
    (the C++ Engine mirror turns the status into std::runtime_error, the reference's convention on this path)."""
    lib = ctypes.CDLL(built["cuda"], mode=ctypes.RTLD_GLOBAL)
    lib.qcm_last_error.restype = ctypes.c_char_p
    assert lib.qcm_init(0) == 0, lib.qcm_last_error()
    assert lib.qcm_init(0) == 0                                    # idempotent on the same device
    arr = ctypes.c_void_p()
    assert lib.qcm_array_alloc(ctypes.c_int64(16), ctypes.byref(arr)) == 0
    host = (ctypes.c_double * 32)(*range(32))
    assert lib.qcm_array_upload(arr, ctypes.c_int64(8), host, ctypes.c_int64(16)) != 0       # 8 + 16 > 16
    assert b"range" in lib.qcm_last_error()
    assert lib.qcm_array_upload(arr, ctypes.c_int64(0), host, ctypes.c_int64(16)) == 0
    back = (ctypes.c_double * 16)()
    assert lib.qcm_array_download(arr, ctypes.c_int64(0), back, ctypes.c_int64(16)) == 0 and list(back) == [float(i) for i in range(16)]
    # a null plan is refused by each of the engine calls
    out = (ctypes.c_double * 16)()
    assert lib.qcm_site_hamil2(None, arr, arr, host, out) != 0 and b"plan" in lib.qcm_last_error()
    assert lib.qcm_boundary_step(None, arr, host, host, arr) != 0 and b"plan" in lib.qcm_last_error()
    assert lib.qcm_hdiag(None, arr, arr, out) != 0 and b"plan" in lib.qcm_last_error()
    # vector algebra checks sizes
    res = ctypes.c_double()
    assert lib.qcm_vec_dot(arr, arr, ctypes.c_int64(32), ctypes.byref(res)) != 0 and b"holds 16" in lib.qcm_last_error()
    assert lib.qcm_vec_dot(arr, arr, ctypes.c_int64(16), ctypes.byref(res)) == 0 and res.value == sum(i * i for i in range(16))
    assert lib.qcm_array_free(arr) == 0
