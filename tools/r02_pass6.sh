#!/bin/bash
# round 2, pass 6 (one GPU): GPU tests including the noise-term tests (tests/test_noise.py), then the cfg3 sigma line without the sweeps
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02i_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02i_pytest_gpu.log
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config-sweep --no-sweep ) > gpurun_out/r02i_bench_cfg3.json 2> gpurun_out/r02i_bench_cfg3.err
grep "rror\|real" gpurun_out/r02i_bench_cfg3.err | tail -4
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02i_bench_cfg3.json"))
    print("cfg3 N=1 value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s frac %.3f roof %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["frac_of_fp64_peak"], d["roofline"]["frac"]))
except Exception as e:
    print("failed", e)
PY
