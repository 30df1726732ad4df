#!/bin/bash
# iteration pass: smoke, parity tests, bench line(s), launch list
mkdir -p gpurun_out
QCM_DEBUG=1 timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -15 gpurun_out/smoke.log
if ! grep -q "smoke 2u1" gpurun_out/smoke.log; then echo "smoke failed, stopping"; exit 1; fi
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 2500 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 600 python bench.py --config cfg3_24e30o_su2u1_M2000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 2500 gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
