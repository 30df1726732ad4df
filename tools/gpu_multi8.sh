#!/bin/bash
# 8-GPU sigma bench at cfg3 only (costly: 8x box time)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n8.json 2> gpurun_out/bench_cfg3_n8.err
grep -v "^\*\|OMP_NUM" gpurun_out/bench_cfg3_n8.err | tail -3 | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/bench_cfg3_n8.json"))
print("cfg3 N=8 value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d["frac_of_fp64_peak"]))
PY
free -g | head -2; nproc
