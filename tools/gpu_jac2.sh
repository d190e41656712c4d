#!/bin/bash
mkdir -p gpurun_out
JAC_CASES="512:1024:4,256:2048:16,128:4096:16" JAC_SCHED="0:0,1:0,1:300,1:600,1:1000,2:0,2:300,2:600,3:0,3:300,4:0" timeout 600 python tools/jacobi_bench.py > gpurun_out/jacsched.log 2>&1
echo rc=$?; cat gpurun_out/jacsched.log | cut -c1-200
