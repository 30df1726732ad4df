#!/bin/bash
# round 2, last N-GPU pass: two-device GPU tests, then the default bench line (cfg3 sigma + oracle parity + cfg3 sweep) on N GPUs with the
# queued boundary steps, the recycled boundary arrays / plan buffers and the corrected two-site sweep order
N=${1:-2}; T=${2:-r02p}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_gpus_n$N.txt; nproc >> gpurun_out/${T}_gpus_n$N.txt
if [ "$N" = "2" ]; then ( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/${T}_pytest_multi_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_multi_gpu.log; fi
( time QCM_DEBUG=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/${T}_bench_cfg3_n$N.json 2> gpurun_out/${T}_bench_cfg3_n$N.err
grep "bench rank 0\|rror\|real\|rank 0\] split" gpurun_out/${T}_bench_cfg3_n$N.err | tail -12 | cut -c1-500
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_bench_cfg3_n$N.json"))
    print("N=$N value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s  parity %s exec %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle"), d["details"]["executed_flops"]))
    s = d.get("config_sweep", {})
    print({k: v for k, v in s.items() if k != "energies"})
except Exception as e:
    print("failed", e)
PY
