#!/bin/bash
# debugging pass: small N-GPU bench with hard timeouts and progress lines
N=${1:-2}; CFG=${2:-cfg2_10e26o_su2u1_M1000}
mkdir -p gpurun_out
( time QCM_SYNC_DEBUG=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config $CFG --steps 5 --warmup 3 --sweep-budget 60 ) > gpurun_out/r02_dbg_n$N.json 2> gpurun_out/r02_dbg_n$N.err
grep "bench rank\|rror\|real" gpurun_out/r02_dbg_n$N.err | tail -30; tail -c 1500 gpurun_out/r02_dbg_n$N.json
