#!/bin/bash
# multi-GPU evidence: bash tools/gpu_r2_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | tail -n +2 | sort | uniq -c > gpurun_out/r2m_gpus_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-roofline > gpurun_out/r2m_bench_n${N}_weak.json 2> gpurun_out/r2m_bench_n${N}_weak.err; echo "weak rc=$?"
head -c 330 gpurun_out/r2m_bench_n${N}_weak.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --no-roofline --scaling strong --global-batch 64 > gpurun_out/r2m_bench_n${N}_strong_config3.json 2> gpurun_out/r2m_bench_n${N}_strong.err; echo "strong rc=$?"
head -c 330 gpurun_out/r2m_bench_n${N}_strong_config3.json; echo
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -k "two_gpu" > gpurun_out/r2m_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_pytest_2gpu.txt
  timeout 600 python bench.py --steps 10 --warmup 3 --no-roofline > gpurun_out/r2m_bench_n1_samebox.json 2> /dev/null; head -c 200 gpurun_out/r2m_bench_n1_samebox.json; echo
  timeout 600 python bench.py --steps 10 --warmup 3 --no-roofline --scaling strong --global-batch 64 > gpurun_out/r2m_bench_n1_strong_config3.json 2> /dev/null; head -c 200 gpurun_out/r2m_bench_n1_strong_config3.json; echo
fi
