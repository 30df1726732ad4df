#!/bin/bash
# N-GPU pass: sharded sigma through torchrun (NCCL allreduce inside the library), parity against the oracle at cfg2
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config cfg2_10e26o_su2u1_M1000 --steps 10 --warmup 3 --parity > gpurun_out/bench_cfg2_n$N.json 2> gpurun_out/bench_cfg2_n$N.err; tail -c 1500 gpurun_out/bench_cfg2_n$N.json; tail -3 gpurun_out/bench_cfg2_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err; tail -c 1500 gpurun_out/bench_cfg3_n$N.json; tail -3 gpurun_out/bench_cfg3_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --config cfg2_10e26o_su2u1_M1000 --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; tail -c 600 gpurun_out/bench_ref_n$N.json; tail -3 gpurun_out/bench_ref_n$N.err
