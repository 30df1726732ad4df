#!/bin/bash
# end-of-round pass: parity tests, the default bench line (cfg3 + CPU baseline + sweep block), reference arm, one ncu --set full
# capture of the dominant kernels (step-1 128x128 GEMM, closing 128x128 GEMM, W class ng=64, W class ng=8)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 700 gpurun_out/bench_reference.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_ws<4, 2, 4, 8|k_wgemm_ws<4, 2, 4, 4|k_wgemm_ws<8, 1" -c 5 -o gpurun_out/prof_cfg3_final -f python tools/profile_sigma.py cfg3_24e30o_su2u1_M2000 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
