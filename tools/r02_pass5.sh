#!/bin/bash
# round 2, pass 5 (one GPU): GPU tests (time-sliced shards, spill tier, descriptor ABI, ...), then cfg4 (24e/30o TwoU1 M=4000) on ONE GPU as two time-sliced shards
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02f_pytest_gpu.log
( time timeout 900 python bench.py --config cfg4_24e30o_2u1_M4000 --steps 3 --warmup 3 --no-cpu-baseline --no-config-sweep --no-sweep ) > gpurun_out/r02f_bench_cfg4_n1.json 2> gpurun_out/r02f_bench_cfg4_n1.err
grep "bench rank 0\|rror\|real" gpurun_out/r02f_bench_cfg4_n1.err | tail -8
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02f_bench_cfg4_n1.json"))
    print("cfg4 N=1 value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s peak %.2f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["fp64_peak_tflops"], d["frac_of_fp64_peak"]))
    print(d["config"])
except Exception as e:
    print("failed", e)
PY
