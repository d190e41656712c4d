#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transform.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2g_pytest.txt 2>&1; tail -2 gpurun_out/r2g_pytest.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -k "config2 or golden or teacher_forced_levels or free_running or batch_equals" > gpurun_out/r2g_pytest2.txt 2>&1; tail -2 gpurun_out/r2g_pytest2.txt
timeout 300 python tools/matfun_bench.py 2>&1 | grep -v "converged at" | head -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> /dev/null; head -c 200 gpurun_out/r2g_bench_n1.json; echo
