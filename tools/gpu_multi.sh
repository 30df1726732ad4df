#!/bin/bash
# N-GPU pass: sharded sigma through torchrun (NCCL allreduce inside the library); parity against the oracle at cfg2, bench at cfg3
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config cfg2_10e26o_su2u1_M1000 --steps 10 --warmup 3 --parity > gpurun_out/bench_cfg2_n$N.json 2> gpurun_out/bench_cfg2_n$N.err; tail -3 gpurun_out/bench_cfg2_n$N.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; tail -3 gpurun_out/bench_cfg3_n$N.err | cut -c1-300
python - <<PY
import json
for c in ("cfg2", "cfg3"):
    try:
        d = json.load(open("gpurun_out/bench_%s_n$N.json" % c))
        print(c, "N=$N value %.2f TF/s  %.2f ms  e2e %.2f TF/s  phases %s  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in d["roofline"]["phase_ms"].items()}, d.get("parity_rel_err_vs_oracle")))
    except Exception as e:
        print(c, "failed", e)
PY
