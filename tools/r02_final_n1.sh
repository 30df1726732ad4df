#!/bin/bash
# round 2, final one-GPU pass: GPU tests, default bench line (cfg3 sigma + sweep), cfg4 on one GPU, reference arm, per-launch ncu counters of one cfg3 sigma
mkdir -p gpurun_out
nproc > gpurun_out/r02g_nproc.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02g_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02g_pytest_gpu.log
( time QCM_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02g_bench_cfg3.json 2> gpurun_out/r02g_bench_cfg3.err
grep "bench rank 0\|rror\|real\|split seconds" gpurun_out/r02g_bench_cfg3.err | tail -10
( time timeout 600 python bench.py --config cfg4_24e30o_2u1_M4000 --steps 3 --warmup 3 --no-cpu-baseline --no-config-sweep --no-sweep ) > gpurun_out/r02g_bench_cfg4_n1.json 2> gpurun_out/r02g_bench_cfg4_n1.err
grep "rror\|real" gpurun_out/r02g_bench_cfg4_n1.err | tail -4
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum
timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r02g_counters_cfg3.csv python tools/profile_sigma.py cfg3_24e30o_su2u1_M2000 1 > gpurun_out/r02g_ncu_counters.log 2>&1
tail -2 gpurun_out/r02g_ncu_counters.log
python - <<PY
import json
for f in ("r02g_bench_cfg3", "r02g_bench_cfg4_n1"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.2f TF/s  %.2f ms  e2e %.2f (engine %.2f) TF/s  phases %s parity %s peak %.2f/%.2f frac %.3f roof %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["engine_mirror"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle"), d["fp64_peak_tflops"], d["fp64_fma_peak_tflops"], d["frac_of_fp64_peak"], d["roofline"]["frac"]))
        s = d.get("config_sweep", {})
        print("   sweep", {k: v for k, v in s.items() if k != "energies"})
        for w in d.get("sweep", []): print("   ", {k: w[k] for k in ("workload", "gpu_seconds_per_sweep", "cpu_seconds_per_sweep", "max_abs_energy_diff_vs_oracle") if k in w})
    except Exception as e:
        print(f, "failed", e)
PY
