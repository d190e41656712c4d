#!/bin/bash
# first GPU contact: staged tests, each under its own timeout, logs to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
run layers_basic python -m pytest tests/test_gpu_layers.py -q -m gpu -k "image or act_roundtrip or conv_head or ref_kernel or maxpool or conv_tail" -x
run conv_tc python -m pytest tests/test_gpu_layers.py -q -m gpu -k "tensor_core"
run transform python -m pytest tests/test_gpu_transform.py -q -m gpu -s
run pipeline python -m pytest tests/test_gpu_pipeline.py -q -m gpu -s
run smoke python __graft_entry__.py --smoke
cat gpurun_out/summary.txt
