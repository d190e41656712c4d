#!/bin/bash
# round-2 GPU call G: generic patch/stride style swap
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transform.py tests/test_gpu_pipeline.py -m gpu -q -s -k "swap or golden or predict_surface" -p no:cacheprovider > gpurun_out/r2g_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.txt; grep -n "style swap C\|final\|passed\|failed\|Error\|error" gpurun_out/r2g_pytest.txt | head -40
