"""BASELINE configs[1] as a real run: N2-sized 10e/26o synthetic FCIDUMP, SU2U1, two-site sweeps truncated to M=1000 on one B200
(qcm/twosite.hpp driver above the GPU engine; random MPS start at M0, the bond dimension grows through the two-site splits).
usage: python tools/gpu_sweep_cfg2.py [nsweeps] [Mmax] [M0]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from qcmaquis_b200 import build
nsweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Mmax = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
M0 = int(sys.argv[3]) if len(sys.argv) > 3 else 64
cu = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL); host = ctypes.CDLL(build.build_host()); host.qcmd_create.restype = ctypes.c_void_p
cu.qcm_launch_count.restype = ctypes.c_int64
path = bench.make_fcidump(26, 10); e = bench.errbuf()
t = time.time()
h = ctypes.c_void_p(host.qcmd_create(path.encode(), b"su2u1", 26, 10, e, 1024))
t_mpo = time.time() - t
en = (ctypes.c_double * 8192)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)()
t = time.time()
rc = host.qcmd_ts_sweeps(h, M0, Mmax, nsweeps, 42, 0, en, 8192, ctypes.byref(n), info, e, 1024)
wall = time.time() - t
per = n.value // max(nsweeps, 1)
out = {"workload": "cfg2_10e26o_su2u1 two-site sweeps, Mmax=%d, random start M0=%d" % (Mmax, M0), "rc": rc, "error": e.value.decode(), "nsweeps": nsweeps,
       "mpo_seconds": t_mpo, "wall_seconds": wall, "sweep_seconds_total": info[1], "sigma_evaluations": int(info[0]), "micro_iterations": n.value,
       "largest_bond_dimension": int(info[3]), "energy_after_each_sweep": [en[(s + 1) * per - 1] for s in range(nsweeps)] if per else [],
       "energies_monotone_after_first_sweep": all(en[i + 1] <= en[i] + 1e-8 for i in range(per, n.value - 1)), "gpu_launches": int(cu.qcm_launch_count())}
print(json.dumps(out))
