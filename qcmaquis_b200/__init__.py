"""qcmaquis_b200 -- B200-native execution of QCMaquis's DMRG sweep hot path (sigma vector + boundary updates).

The product is native: CUDA kernels and a C ABI (include/qcm_b200.h, qcmaquis_b200/lib/libqcm_b200.so) under a
C++ host that mirrors contraction::Engine (qcmaquis_b200/csrc/qcm/engine_gpu.hpp).  This Python package only
builds the libraries in-tree and loads them for the tests and bench.py."""
import ctypes, os
from . import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cuda():
    """The C ABI library. Raises if it cannot be built/loaded; there is no fallback implementation."""
    lib = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL)
    lib.qcm_last_error.restype = ctypes.c_char_p
    lib.qcm_stream.restype = ctypes.c_void_p
    lib.qcm_launch_count.restype = ctypes.c_int64
    lib.qcm_array_devptr.restype = ctypes.c_void_p
    return lib


def load_host():
    load_cuda()
    lib = ctypes.CDLL(build.build_host())
    lib.qcmd_create.restype = ctypes.c_void_p
    lib.qcmd_plan_flops.restype = ctypes.c_double
    return lib
