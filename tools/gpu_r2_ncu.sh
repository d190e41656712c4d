#!/bin/bash
# round-2 evidence: ncu captures (1 GPU).  Raw reports land in gpurun_out/, tools/summarize_ncu_r2.py writes profiles/r02_*.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py 2 > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2 -o gpurun_out/r2_conv_b30 -f python tools/ncu_conv.py 30 > gpurun_out/r2_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_jacobi -o gpurun_out/r2_jacobi -f python tools/jacobi_once.py > gpurun_out/r2_ncu_jacobi.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:cov_tc_kernel -o gpurun_out/r2_cov -f python tools/cov_once.py 16 > gpurun_out/r2_ncu_cov.log 2>&1
ls -la gpurun_out/r2_*
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_(tail|head)_tc_kernel" -s 2 -c 2 -o gpurun_out/r2_tail -f python tools/tail_head_once.py 16 > gpurun_out/r2_ncu_tail.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns_gemm -s 55 -c 1 -o gpurun_out/r2_nsgemm -f python tools/matfun_once.py > gpurun_out/r2_ncu_nsgemm.log 2>&1
ls -la gpurun_out/r2_tail* gpurun_out/r2_nsgemm*
