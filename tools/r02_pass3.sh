#!/bin/bash
# round 2, pass 3 (two GPUs): GPU tests (descriptor ABI, sweeps, multi-GPU), then the cfg3 two-site sweep on 2 GPUs with the split timers
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02c_pytest_gpu.log
( time QCM_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --sweep-budget 240 --no-cpu-baseline ) > gpurun_out/r02c_bench_cfg3_n2.json 2> gpurun_out/r02c_bench_cfg3_n2.err
grep "bench rank 0\|rror\|real\|split seconds" gpurun_out/r02c_bench_cfg3_n2.err | tail -12
python - <<PY
import json
for f in ("r02c_bench_cfg3_n2",):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}))
        s = d.get("config_sweep", {})
        print("   sweep", {k: v for k, v in s.items() if k != "energies"})
    except Exception as e:
        print(f, "failed", e)
PY
