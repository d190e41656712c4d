#!/bin/bash
# 2-GPU run: torchrun bench (frame-sharded, NCCL gather) + N=1 for the scaling ratio
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo n1 rc=$?
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1; echo n$N rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.log 2>&1; echo ref rc=$?
tail -1 gpurun_out/bench_n1.log | cut -c1-160; tail -1 gpurun_out/bench_n$N.log | cut -c1-400; tail -2 gpurun_out/bench_ref_n$N.log | cut -c1-200
