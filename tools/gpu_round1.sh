#!/bin/bash
# first GPU pass of a round: parity tests, peak probes, bench lines, launch list, one full ncu capture per hot kernel
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; nvidia-smi -L > gpurun_out/gpus.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_probe.py > gpurun_out/peaks.json 2> gpurun_out/peaks.err; cat gpurun_out/peaks.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json
timeout 600 python bench.py --config cfg3_24e30o_su2u1_M2000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -c 3000 gpurun_out/bench_cfg3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 2 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_dmma -s 4 -c 4 -o gpurun_out/prof_gemm -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wapply -c 2 -o gpurun_out/prof_wapply -f python tools/profile_sigma.py cfg2_10e26o_su2u1_M1000 1 > gpurun_out/ncu_wapply.log 2>&1
ls -la gpurun_out
