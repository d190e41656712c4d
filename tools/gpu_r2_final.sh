#!/bin/bash
# round-2 final evidence run (1 GPU): full GPU test-suite with parity printout, bench (+ reference arm), per-call breakdown,
# other configs, per-layer conv bench, eigensolver bench, covariance bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2z_gpu.txt; nproc >> gpurun_out/r2z_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2z_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.txt; tail -4 gpurun_out/r2z_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> /dev/null; echo "ref rc=$?"
timeout 300 python tools/step_breakdown.py 16 > gpurun_out/r2z_step_breakdown_batch16.txt 2>&1
timeout 300 python tools/config_runs.py > gpurun_out/r2z_other_configs.json 2>&1
timeout 300 python tools/conv_bench.py 16 > gpurun_out/r2z_conv_layer_bench.txt 2>&1
timeout 300 python tools/jacobi_bench.py > gpurun_out/r2z_jacobi_bench.txt 2>&1
timeout 300 python tools/cov_bench.py 16 > gpurun_out/r2z_cov_bench.txt 2>&1
timeout 300 python tests/noise_split_gpu.py 128 > gpurun_out/r2z_noise_gpu.txt 2>&1
timeout 300 python tools/tail_bench.py 16 > gpurun_out/r2z_tail_bench.txt 2>&1
timeout 300 python tools/head_bench.py 16 > gpurun_out/r2z_head_bench.txt 2>&1
timeout 300 python tools/image_ops_bench.py > gpurun_out/r2z_image_ops_bench.txt 2>&1
timeout 300 python tools/timeline.py 30 2 1 > gpurun_out/r2z_timeline.txt 2>&1
timeout 300 python tools/matfun_bench.py > gpurun_out/r2z_matfun_bench.txt 2>&1
timeout 300 python tools/k_probe.py > gpurun_out/r2z_k_probe.txt 2>&1
head -c 400 gpurun_out/r2z_bench_n1.json; echo; cat gpurun_out/r2z_other_configs.json | tail -12
