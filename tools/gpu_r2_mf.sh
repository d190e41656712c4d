#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -p no:cacheprovider -s > gpurun_out/r2f_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt; tail -12 gpurun_out/r2f_pytest.txt
timeout 1500 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -s > gpurun_out/r2f_pytest2.txt 2>&1
echo "pytest2 rc=$?" >> gpurun_out/r2f_pytest2.txt; tail -8 gpurun_out/r2f_pytest2.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2f_bench_n1.json; echo
