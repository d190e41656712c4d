#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transform.py tests/test_gpu_pipeline.py -q -m gpu -x -s 2>&1 | grep -E "null value|passed|failed|Error|assert" | head
JAC_CASES="512:1024:4,512:300:4,512:100:4,256:200:8,128:4096:8,64:40:8" timeout 300 python tools/jacobi_bench.py 2>&1 | tee gpurun_out/jac_rankdef.log | cut -c1-200
timeout 200 python tools/rank_probe.py 2>&1 | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
