#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_jacobi" -c 1 -f -o gpurun_out/prof_jac python tools/profile_step.py 2 > gpurun_out/ncu_jac.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_jac.log
