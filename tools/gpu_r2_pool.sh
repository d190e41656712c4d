#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -p no:cacheprovider -k "pool2" > gpurun_out/r2p_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.txt; tail -6 gpurun_out/r2p_pytest.txt
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -k "pool_fused or golden or teacher_forced_levels" > gpurun_out/r2p_pytest2.txt 2>&1
echo "pytest2 rc=$?" >> gpurun_out/r2p_pytest2.txt; tail -4 gpurun_out/r2p_pytest2.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2p_bench_n1.json; echo
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
k=d['kernel_ms_per_step']
for a,b in sorted(k.items(), key=lambda x:-x[1]):
    if 'pool' in a or '64x64' in a: print("%-34s %7.3f"%(a,b))
PY
