#!/bin/bash
# iteration pass: smoke, parity tests, bench line(s), per-launch counters of one sigma for cfg2 and cfg3
mkdir -p gpurun_out
QCM_DEBUG=1 timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep -v "variant\|class" gpurun_out/smoke.log | tail -5
if ! grep -q "smoke 2u1" gpurun_out/smoke.log; then echo "smoke failed, stopping"; exit 1; fi
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config cfg2_10e26o_su2u1_M1000 --steps 10 --warmup 3 --no-sweep > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 2500 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 2500 gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_cfg3.err
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum
for c in cfg3_24e30o_su2u1_M2000; do
  QCM_DEBUG=1 timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/counters_$c.csv python tools/profile_sigma.py $c 1 > gpurun_out/ncu_counters_$c.log 2>&1
  tail -1 gpurun_out/ncu_counters_$c.log
done
