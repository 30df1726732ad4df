#!/bin/bash
# round 2, pass 4 (one GPU): all GPU tests, default bench line, launch list of one sigma
mkdir -p gpurun_out
nproc > gpurun_out/r02e_nproc.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02e_pytest_gpu.log
( time QCM_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02e_bench_cfg3.json 2> gpurun_out/r02e_bench_cfg3.err
grep "bench rank 0\|rror\|real\|split seconds" gpurun_out/r02e_bench_cfg3.err | tail -12
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02e_bench_cfg3.json"))
    print("N=1 value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s parity %s peak %.2f frac %.3f roof %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle"), d["fp64_peak_tflops"], d["frac_of_fp64_peak"], d["roofline"]["frac"]))
    s = d.get("config_sweep", {})
    print("   sweep", {k: v for k, v in s.items() if k != "energies"})
    for w in d.get("sweep", []): print("   ", {k: w[k] for k in ("workload", "gpu_seconds_per_sweep", "cpu_seconds_per_sweep", "max_abs_energy_diff_vs_oracle") if k in w})
except Exception as e:
    print("failed", e)
PY
