#!/bin/bash
# round-2 GPU call F (2 GPUs): sharded == single bit-for-bit, weak- and strong-scaling bench lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpus.txt
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -s -k "config3 or batch_equals or grouped" -p no:cacheprovider > gpurun_out/r2f_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt; tail -5 gpurun_out/r2f_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "weak rc=$?"; head -c 700 gpurun_out/r2f_bench_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong --global-batch 64 --no-roofline > gpurun_out/r2f_bench_n2_strong.json 2> gpurun_out/r2f_bench_n2_strong.err
echo "strong rc=$?"; head -c 900 gpurun_out/r2f_bench_n2_strong.json; echo; tail -3 gpurun_out/r2f_bench_n2_strong.err
timeout 300 python bench.py --steps 5 --warmup 3 --scaling strong --global-batch 64 --no-roofline --no-cpu-baseline > gpurun_out/r2f_bench_n1_strong.json 2> gpurun_out/r2f_bench_n1_strong.err
echo "strong n1 rc=$?"; head -c 300 gpurun_out/r2f_bench_n1_strong.json; echo
