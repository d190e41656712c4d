#!/bin/bash
# round-2 GPU call B: layer tests after a conv-kernel change, per-layer conv bench, encoder noise split, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_transform.py -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/r2b_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt; tail -3 gpurun_out/r2b_pytest.txt
timeout 300 python tools/conv_bench.py 16 > gpurun_out/r2b_convbench.txt 2>&1; cat gpurun_out/r2b_convbench.txt
timeout 300 python tests/noise_split_gpu.py 128 > gpurun_out/r2b_noise_gpu.txt 2>&1; cat gpurun_out/r2b_noise_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['share_of_step'])
print(json.dumps(d['kernel_ms_per_step']))
print(json.dumps(d['covariance_hbm']))
PY
