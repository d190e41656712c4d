#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transform.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python tools/step_breakdown.py 16 2>&1 | grep -E "serial|wct_|overlapped"
