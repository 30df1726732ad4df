#!/bin/bash
# ncu --set full of launches 30..34 of one cfg3 sigma: W class ng=8, k_wstream, closing GEMM 128x128 / 64x128 / 128x64
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 30 --launch-count 5 -o gpurun_out/prof_cfg3_final -f python tools/profile_sigma.py cfg3_24e30o_su2u1_M2000 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
