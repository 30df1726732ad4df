#!/bin/bash
# round 2: BASELINE config 5 (54e/54o SU2U1 M=3000, centre two-site problem) on ONE B200 as time-sliced shards of the MPO bond graph
mkdir -p gpurun_out
S=${1:-8}
free -g | head -2 > gpurun_out/r02m_cfg5_mem.txt
( time QCM_DEBUG= timeout 1000 python bench.py --config cfg5_54e54o_su2u1_M3000 --slices $S --steps 2 --warmup 3 --no-cpu-baseline --no-config-sweep --no-sweep ) > gpurun_out/r02m_bench_cfg5_n1.json 2> gpurun_out/r02m_bench_cfg5_n1.err
grep "bench rank 0\|rror\|real\|Killed" gpurun_out/r02m_bench_cfg5_n1.err | tail -12
nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02m_bench_cfg5_n1.json"))
    print("cfg5 N=1 value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s peak %.2f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["fp64_peak_tflops"], d["frac_of_fp64_peak"]))
    print(d["config"], d["details"])
except Exception as e:
    print("failed", e)
PY
