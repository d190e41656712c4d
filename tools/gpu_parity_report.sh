#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transform.py tests/test_gpu_pipeline.py -q -m gpu -s > gpurun_out/parity.log 2>&1; echo parity rc=$?
grep -E "free-running|encoder .* transform|max-abs vs reference|sweeps|passed|failed" gpurun_out/parity.log | head -60
timeout 600 python tools/config_runs.py > gpurun_out/configs.log 2>&1; echo cfg rc=$?; tail -12 gpurun_out/configs.log
