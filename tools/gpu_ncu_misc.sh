#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_conv_head|k_conv_tail|k_cov_partial|k_chan_sums|k_jacobi" -c 14 -f -o gpurun_out/prof_misc python tools/profile_step.py 2 > gpurun_out/ncu_misc.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_misc.log
