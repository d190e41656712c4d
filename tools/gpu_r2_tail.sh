#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -p no:cacheprovider -k "tail" > gpurun_out/r2t_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.txt; tail -12 gpurun_out/r2t_pytest.txt
timeout 300 python tools/tail_bench.py 16 > gpurun_out/r2t_tail_bench.txt 2>&1; cat gpurun_out/r2t_tail_bench.txt
