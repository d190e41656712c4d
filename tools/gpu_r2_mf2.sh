#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/matfun_bench.py > gpurun_out/r2f_matfun_bench.txt 2>&1; grep -v "converged at" gpurun_out/r2f_matfun_bench.txt | head -20
timeout 1200 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2f_pytest.txt 2>&1; tail -2 gpurun_out/r2f_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"; head -c 200 gpurun_out/r2f_bench_n1.json; echo
