#!/bin/bash
# round 2, eight GPUs: the default bench line (cfg3 sigma with oracle parity + cfg3 two-site sweep) on 8 B200
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus_n$N.txt; nproc >> gpurun_out/r02_gpus_n$N.txt
( time QCM_DEBUG=1 timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 --sweep-budget 200 ) > gpurun_out/r02d_bench_cfg3_n$N.json 2> gpurun_out/r02d_bench_cfg3_n$N.err
grep "bench rank 0\|rror\|real\|rank 0\] split" gpurun_out/r02d_bench_cfg3_n$N.err | tail -14
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02d_bench_cfg3_n$N.json"))
    print("N=$N value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle")))
    print("   exec", d["details"]["executed_flops"], "frac", d["frac_of_fp64_peak"])
    s = d.get("config_sweep", {})
    print("   sweep", {k: v for k, v in s.items() if k != "energies"})
except Exception as e:
    print("failed", e)
PY
