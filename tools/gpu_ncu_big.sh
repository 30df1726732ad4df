#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -s 4 -c 1 -o gpurun_out/prof_gemm_step1 -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_step1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 24 -c 1 -o gpurun_out/prof_gemm_step3 -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_step3.log 2>&1
./tools/probes/dmma_probe > gpurun_out/dmma_probe.txt 2>&1
tail -3 gpurun_out/ncu_step1.log gpurun_out/ncu_step3.log; cat gpurun_out/dmma_probe.txt
