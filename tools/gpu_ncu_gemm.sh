#!/bin/bash
# per-launch counters of every kernel of one sigma evaluation (cfg2) + the plan's per-launch FLOPs, then one full capture
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum
QCM_DEBUG=1 timeout 900 ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/counters_cfg2.csv python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_counters.log 2>&1
tail -2 gpurun_out/ncu_counters.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -c 2 -o gpurun_out/prof_gemm_ws -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_gemm_ws.log 2>&1
tail -2 gpurun_out/ncu_gemm_ws.log; ls -la gpurun_out/
