#!/bin/bash
# one GPU: the cfg3 two-site sweep only (no CPU baseline, no small sweeps)
mkdir -p gpurun_out
( time QCM_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sweep ) > gpurun_out/r02h_bench_cfg3.json 2> gpurun_out/r02h_bench_cfg3.err
grep "bench rank 0\|rror\|real\|split seconds" gpurun_out/r02h_bench_cfg3.err | tail -8
python - <<PY
import json
d = json.load(open("gpurun_out/r02h_bench_cfg3.json"))
s = d.get("config_sweep", {})
print("sigma %.2f ms; sweep" % d["ms_per_step"], {k: v for k, v in s.items() if k != "energies"})
PY
