#!/bin/bash
# round 2, pass 7 (one GPU): ragged tiles issue only the fragments inside the block (FRAG_SKIP), plan creation on several host threads.
# GPU tests, then the cfg3 sigma line with oracle parity (no sweeps), with the per-variant useful FLOPs (QCM_DEBUG)
mkdir -p gpurun_out
T=${1:-r02j}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -6 gpurun_out/${T}_pytest_gpu.log
( time QCM_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-config-sweep --no-sweep ) > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err
grep "rror\|real\|bench rank" gpurun_out/${T}_bench_cfg3.err | tail -12
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_cfg3.json"))
    print("cfg3 N=1 value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s frac %.3f roof %.3f parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["frac_of_fp64_peak"], d["roofline"]["frac"], d.get("parity_rel_err_vs_oracle")))
except Exception as e:
    print("failed", e)
PY
