#!/bin/bash
# round 2, pass 2 (two GPUs): all GPU tests, then a bounded 2-GPU bench at cfg2 (sigma + sweep) and at cfg3 (sigma + sweep)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02b_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02b_pytest_gpu.log
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config cfg2_10e26o_su2u1_M1000 --steps 5 --warmup 3 --sweep-budget 120 ) > gpurun_out/r02b_bench_cfg2_n2.json 2> gpurun_out/r02b_bench_cfg2_n2.err
grep "bench rank 0\|rror\|real" gpurun_out/r02b_bench_cfg2_n2.err | tail -12
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --sweep-budget 240 ) > gpurun_out/r02b_bench_cfg3_n2.json 2> gpurun_out/r02b_bench_cfg3_n2.err
grep "bench rank 0\|rror\|real" gpurun_out/r02b_bench_cfg3_n2.err | tail -12
python - <<PY
import json
for f in ("r02b_bench_cfg2_n2", "r02b_bench_cfg3_n2"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle")))
        print("   exec", d["details"]["executed_flops"])
        s = d.get("config_sweep", {})
        print("   sweep", {k: v for k, v in s.items() if k != "energies"})
    except Exception as e:
        print(f, "failed", e)
PY
