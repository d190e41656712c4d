#!/bin/bash
# scope row 8f-3 (device image ops): parity tests + a timing line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_image_ops.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2i_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.txt; tail -15 gpurun_out/r2i_pytest.txt
timeout 300 python tools/image_ops_bench.py > gpurun_out/r2i_image_ops_bench.txt 2>&1; cat gpurun_out/r2i_image_ops_bench.txt
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_transform.py -m gpu -q -x -p no:cacheprovider -k "cli or video or passes or predict_surface or jacobi" > gpurun_out/r2i_pytest2.txt 2>&1
echo "pytest2 rc=$?" >> gpurun_out/r2i_pytest2.txt; tail -5 gpurun_out/r2i_pytest2.txt
