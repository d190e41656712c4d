#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py -q -m gpu -x -k "tensor_core" > gpurun_out/tests_conv.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_conv.log
timeout 600 python tools/conv_bench.py ${BATCH:-8} > gpurun_out/conv_bench.log 2>&1; echo rc=$?
cat gpurun_out/conv_bench.log
