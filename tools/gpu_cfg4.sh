#!/bin/bash
# abelian path at scale: cfg4 24e/30o TwoU1 M=4000 needs 145 GB of resident step-1 products -> two or more B200
N=${1:-2}
mkdir -p gpurun_out
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --config cfg4_24e30o_2u1_M4000 --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err; tail -c 2500 gpurun_out/bench_cfg4_n$N.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench_cfg4_n$N.err | tail -4 | cut -c1-400
