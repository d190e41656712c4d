#!/bin/bash
mkdir -p gpurun_out
for q in 1e-4 3e-4 1e-3; do
  echo "== tolq $q"
  JAC_TOLQ=$q JAC_CASES="512:1024:4,256:2048:8,128:4096:8,64:4096:8,512:300:4" timeout 200 python tools/jacobi_bench.py 2>&1 | cut -c1-175
done | tee gpurun_out/jac_tolq.log
