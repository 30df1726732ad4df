"""Where a GPU sweep spends its host time: runs one two-site sweep (12e/12o SU2U1 M=300) with QCM_PLAN_TIMING and sums the
planner phases printed on stderr; compares with the sweep's wall time.  usage: QCM_PLAN_TIMING=1 python tools/gpu_sweep_profile.py 2> log"""
import ctypes, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from qcmaquis_b200 import build
cu = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL); host = ctypes.CDLL(build.build_host()); host.qcmd_create.restype = ctypes.c_void_p
norb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
M = int(sys.argv[2]) if len(sys.argv) > 2 else 300
kind = sys.argv[3] if len(sys.argv) > 3 else "ts"
path = bench.make_fcidump(norb, norb); e = bench.errbuf()
h = ctypes.c_void_p(host.qcmd_create(path.encode(), b"su2u1", norb, norb, e, 1024))
en = (ctypes.c_double * 4096)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)()
t = time.time()
if kind == "ts":
    rc = host.qcmd_ts_sweeps(h, 60, M, 1, 42, 0, en, 4096, ctypes.byref(n), info, e, 1024)
else:
    rc = host.qcmd_ss_sweeps(h, M, 1, 42, 0, en, 4096, ctypes.byref(n), info, e, 1024)
print("%de%do su2u1 M=%d %s sweep on the GPU engine: rc %d %s wall %.2f s, sweep %.2f s, micro-iterations %d, sigma %d, final energy %.12f" %
      (norb, norb, M, kind, rc, e.value, time.time() - t, info[1], n.value, int(info[0]), info[2]))
