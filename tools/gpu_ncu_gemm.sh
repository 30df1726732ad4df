#!/bin/bash
# one ncu --set full capture of the GEMM launches of one sigma evaluation (cfg2): step-1 (2 launches) and closing GEMM
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_ws -c 3 -o gpurun_out/prof_gemm_ws -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_gemm_ws.log 2>&1
tail -3 gpurun_out/ncu_gemm_ws.log; ls -la gpurun_out/*.ncu-rep
