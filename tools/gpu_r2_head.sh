#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -p no:cacheprovider -k "head or tail" > gpurun_out/r2h_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.txt; tail -12 gpurun_out/r2h_pytest.txt
timeout 300 python tools/head_bench.py 16 > gpurun_out/r2h_head_bench.txt 2>&1; cat gpurun_out/r2h_head_bench.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2h_bench_n1.json
