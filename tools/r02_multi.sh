#!/bin/bash
# round 2, N-GPU pass: GPU tests that need two devices, then the default bench line (cfg3 sigma + oracle parity + cfg3 sweep) on N GPUs
N=${1:-2}; EXTRA="${2:-}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpus_n$N.txt; nproc >> gpurun_out/r02_gpus_n$N.txt
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r02_pytest_multi_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_multi_gpu.log; fi
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA ) > gpurun_out/r02_bench_cfg3_n$N.json 2> gpurun_out/r02_bench_cfg3_n$N.err; tail -4 gpurun_out/r02_bench_cfg3_n$N.err | cut -c1-400
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_cfg3_n$N.json"))
    print("N=$N value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s  parity %s exec %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle"), d["details"]["executed_flops"]))
    s = d.get("config_sweep", {})
    print({k: v for k, v in s.items() if k != "energies"})
except Exception as e:
    print("failed", e)
PY
