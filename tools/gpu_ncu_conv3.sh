#!/bin/bash
mkdir -p gpurun_out
CONV_BENCH_ONLY=0,6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 4 -c 2 -f -o gpurun_out/prof_conv3 python tools/conv_bench.py 4 > gpurun_out/ncu_conv3.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_conv3.log
