#!/bin/bash
# round 2, pass 10 (one GPU): boundary arrays and plan task buffers recycled (no device-pool allocation per site).  GPU tests, then the cfg3
# sigma line with the cfg3 sweep only (no CPU baseline, no small sweeps), QCM_DEBUG host-time breakdown
mkdir -p gpurun_out
T=${1:-r02o}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -6 gpurun_out/${T}_pytest_gpu.log
( time QCM_DEBUG=1 timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sweep ) > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
grep "bench rank 0\|rror\|real\|split seconds" gpurun_out/${T}_bench_cfg3.err | tail -12
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_cfg3.json"))
    print("value %.2f TF/s  %.2f ms  e2e %.2f (engine %.2f) TF/s  phases %s frac %.3f roof %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["engine_mirror"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["frac_of_fp64_peak"], d["roofline"]["frac"]))
    s = d.get("config_sweep", {})
    print("   sweep", {k: v for k, v in s.items() if k != "energies"})
except Exception as e:
    print("failed", e)
PY
