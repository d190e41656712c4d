#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -p no:cacheprovider -k "pool2" > gpurun_out/r2p_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.txt; tail -6 gpurun_out/r2p_pytest.txt
timeout 1500 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2p_pytest2.txt 2>&1
echo "pytest2 rc=$?" >> gpurun_out/r2p_pytest2.txt; tail -8 gpurun_out/r2p_pytest2.txt
