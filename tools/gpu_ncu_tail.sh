#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_conv_tail_tiled" -c 1 -f -o gpurun_out/prof_tail python tools/profile_step.py 8 > gpurun_out/ncu_tail.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_tail.log
