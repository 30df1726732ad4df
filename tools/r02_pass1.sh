#!/bin/bash
# round 2, pass 1: smoke, GPU tests, default bench line (cfg3 sigma + cfg3 two-site sweep) on one B200
mkdir -p gpurun_out
nproc > gpurun_out/r02_nproc.txt; free -g | head -2 >> gpurun_out/r02_nproc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu.log
( time QCM_DEBUG=1 timeout 1500 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
tail -c 6000 gpurun_out/r02_bench_cfg3.json; grep -v "variant\|gemm launch" gpurun_out/r02_bench_cfg3.err | tail -15
