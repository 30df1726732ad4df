"""Minimal driver for ncu: set up one synthetic site problem and run N device-resident sigma evaluations.
usage: python tools/profile_sigma.py <config> <n_sigma> [site]"""
import ctypes, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench
from qcmaquis_b200 import build
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_10e26o_su2u1_M1000"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
norb, nelec, symm, M = bench.CONFIGS[cfg]
site = int(sys.argv[3]) if len(sys.argv) > 3 else norb // 2 - 1
cu = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL)
host = ctypes.CDLL(build.build_host())
cu.qcm_last_error.restype = ctypes.c_char_p
host.qcmd_create.restype = ctypes.c_void_p
assert cu.qcm_init(0) == 0, cu.qcm_last_error()
e = ctypes.create_string_buffer(1024)
path = bench.make_fcidump(norb, nelec)
h = ctypes.c_void_p(host.qcmd_create(path.encode(), symm.encode(), norb, nelec, e, 1024))
assert h.value, e.value
info = (ctypes.c_double * 32)()
assert host.qcmd_setup_site(h, site, 1, M, 1, 0, 0, 1, info, e, 1024) == 0, e.value
print("flops %.4e launches/sigma %d waves %d" % (info[0], info[20], info[11]), flush=True)
assert host.qcmd_sigma_dev(h, n, e, 1024) == 0, e.value
cu.qcm_sync()
print("done")
