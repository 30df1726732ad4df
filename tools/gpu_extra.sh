#!/bin/bash
# one-off measurements: abelian path at scale (cfg4 24e/30o TwoU1 M=4000) and a 16e/16o M=400 single-site sweep on the GPU engine
mkdir -p gpurun_out
timeout 1500 python bench.py --config cfg4_24e30o_2u1_M4000 --steps 3 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 1800 gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err
timeout 900 python - <<'PY' > gpurun_out/sweep_16o.log 2>&1
import ctypes, os, sys, time
sys.path.insert(0, os.getcwd())
import bench
from qcmaquis_b200 import build
cu = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL); host = ctypes.CDLL(build.build_host()); host.qcmd_create.restype = ctypes.c_void_p
path = bench.make_fcidump(16, 16); e = bench.errbuf()
h = ctypes.c_void_p(host.qcmd_create(path.encode(), b"su2u1", 16, 16, e, 1024))
en = (ctypes.c_double * 4096)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)()
t = time.time(); rc = host.qcmd_ss_sweeps(h, 400, 1, 42, 0, en, 4096, ctypes.byref(n), info, e, 1024)
print("16e16o su2u1 M=400, 1 single-site sweep on the GPU engine: rc", rc, e.value, "wall %.2f s" % (time.time() - t), "micro-iterations", n.value, "sigma", info[0], "sweep seconds", info[1], "final energy %.12f" % info[2])
PY
cat gpurun_out/sweep_16o.log | tail -3
