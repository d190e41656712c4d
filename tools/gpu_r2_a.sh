#!/bin/bash
# round-2 GPU call A: full GPU test-suite (with the at-size parity tests), bench line, per-layer conv bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
nproc >> gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=10 -p no:cacheprovider > gpurun_out/r2a_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; cat gpurun_out/r2a_bench.json | head -c 3000
timeout 300 python tools/conv_bench.py 16 > gpurun_out/r2a_convbench.txt 2>&1
cat gpurun_out/r2a_convbench.txt
