#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2 -o gpurun_out/r2c_conv python tools/ncu_conv.py 4 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
timeout 300 python tools/conv_bench.py 16 > gpurun_out/r2c_convbench.txt 2>&1; cat gpurun_out/r2c_convbench.txt
